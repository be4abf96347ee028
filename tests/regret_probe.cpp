// regret_probe.cpp — test helper: exposes the PRODUCT's host-side regret table builder (urlearning-cpp_b200/csrc/regret.hpp, plain
// C++, no CUDA) through a C entry point so that a CPU test can compare it with the oracle's independent restatement.
#include <cstdint>
#include "../urlearning-cpp_b200/csrc/regret.hpp"

extern "C" void probe_log_regret(int64_t n_max, int r, float *out) {
    const std::vector<float> t = urlgpu::regret::log_regret(n_max, r);
    for (int64_t i = 0; i <= n_max; i++) out[i] = t[(size_t)i];
}
