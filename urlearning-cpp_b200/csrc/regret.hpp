// regret.hpp — host side of the fNML score: the multinomial NML regret table of the reference.
//
//   reg2 / reg / getRegretCache      scoring_function/fnml_scoring_function.h (r2_1000 literals, :reg2, :reg, :getRegretCache)
//
// fNML(v | S) = LL(v | S) - sum_j log C(N_ij, r_v)   (fnml_scoring_function.cpp:28-74, no BIC penalty), with
// C(N, 1) = 1, C(N, 2) = the K=2 regret (tabulated exactly for N <= 1000, Szpankowski's approximation above) and the
// linear recurrence C(N, k) = C(N, k-1) + N/(k-2) C(N, k-2), evaluated in float32 in the reference's operation order.
// The table entry is logf() of that float — the reference's `regret->at(r)->at(N)`.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace urlgpu {
namespace regret {

static const uint32_t kR2Bits[1001] = {
#include "regret_r2.inc"
};

inline float r2(int N) {
    if (N <= 1000) { float f; memcpy(&f, &kR2Bits[N], 4); return f; }
    const double pi = 3.1415926535897932384626433832795;
    // exp(0.5 * log(N * pi / 2) + sqrt(8 / (9 * N * pi)) + (1.0 / 12 - 4 / (9 * pi)) / N), integer sub-expressions as written there
    volatile double a = 0.5 * std::log(N * pi / 2);
    volatile double b = std::sqrt(8 / (9 * N * pi));
    volatile double c = (1.0 / 12 - 4 / (9 * pi)) / N;
    volatile double s = a + b;
    s = s + c;
    return (float)std::exp(s);
}

// volatile float temporaries: no FMA contraction and no excess precision, whatever the host compiler flags
inline float reg(int N, int K) {
    if (K == 1) return 1.0f;
    if (K == 2) return r2(N);
    volatile float rk_2 = 1.0f, rk_1 = r2(N), rk = 0;
    for (int k = 3; k <= K; ++k) {
        volatile float q = rk_2 / (float)(k - 2);
        volatile float m = q * (float)N;
        rk = rk_1 + m;
        rk_2 = rk_1;
        rk_1 = rk;
    }
    return rk;
}

// out[N] = (float)log(reg(N, r)) for N = 0..n_max.  r = 0 is never looked up (an arity is at least 1).
inline std::vector<float> log_regret(int64_t n_max, int r) {
    std::vector<float> out((size_t)n_max + 1);
    // log of a float: the float overload (logf), as in the reference (`log(scoring::reg(n, r))` under <math.h>)
    for (int64_t N = 0; N <= n_max; N++) { const float x = reg((int)N, r); out[(size_t)N] = std::log(x); }
    return out;
}

} // namespace regret
} // namespace urlgpu
