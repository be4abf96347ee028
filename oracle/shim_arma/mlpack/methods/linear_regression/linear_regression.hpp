// shim_arma/.../linear_regression.hpp — mlpack 3.x regression::LinearRegression without intercept and weights, from its
// published algorithm: Train solves (P P^T + lambda I) b = P r^T, ComputeError is the mean squared residual.  The arithmetic
// order is the oracle's (oracle.cpp orc_cbic_the_score_residual: plain dot products, Gaussian elimination with partial
// pivoting), so the REFERENCE'S code compiled over this shim and the oracle agree bit for bit on the_score — what is pinned
// is everything BIC_OLS.cpp does around the solve.  TEST INFRASTRUCTURE.
#pragma once
#include <armadillo>
#include <cmath>
#include <utility>
namespace mlpack { namespace regression {
class LinearRegression {
public:
    LinearRegression(const arma::mat &predictors /*k x n*/, const arma::rowvec &responses, double lambda = 0, bool intercept = true) {
        if (intercept) throw std::logic_error("shim LinearRegression: intercept not supported");
        const int k = (int)predictors.n_rows;
        const long n = (long)predictors.n_cols;
        std::vector<double> cov((size_t)k * k), rhs(k);
        for (int a = 0; a < k; a++) {
            for (int b = a; b < k; b++) {
                double s = 0;
                for (long r = 0; r < n; r++) s += predictors(a, r) * predictors(b, r);
                cov[a * k + b] = cov[b * k + a] = s;
            }
            double s = 0;
            for (long r = 0; r < n; r++) s += predictors(a, r) * responses(r);
            rhs[a] = s;
        }
        for (int a = 0; a < k; a++) cov[a * k + a] += lambda;
        ok = solve(cov, rhs, k);
        parameters = arma::vec(k);
        for (int a = 0; a < k; a++) parameters(a) = ok ? rhs[a] : std::nan("");
    }
    const arma::vec &Parameters() const { return parameters; }
    bool Intercept() const { return false; }   // only printed by the reference
    double ComputeError(const arma::mat &points, const arma::rowvec &responses) const {
        const int k = (int)points.n_rows;
        const long n = (long)points.n_cols;
        double cost = 0;
        for (long r = 0; r < n; r++) {
            double pred = 0;
            for (int a = 0; a < k; a++) pred += parameters(a) * points(a, r);
            const double t = responses(r) - pred;
            cost += t * t;
        }
        return cost / (double)n;
    }
private:
    static bool solve(std::vector<double> &a, std::vector<double> &b, int k) {
        for (int c = 0; c < k; c++) {
            int piv = c;
            for (int r = c + 1; r < k; r++) if (std::fabs(a[r * k + c]) > std::fabs(a[piv * k + c])) piv = r;
            if (a[piv * k + c] == 0) return false;
            if (piv != c) { for (int j = 0; j < k; j++) std::swap(a[c * k + j], a[piv * k + j]); std::swap(b[c], b[piv]); }
            for (int r = c + 1; r < k; r++) {
                const double f = a[r * k + c] / a[c * k + c];
                for (int j = c; j < k; j++) a[r * k + j] -= f * a[c * k + j];
                b[r] -= f * b[c];
            }
        }
        for (int r = k - 1; r >= 0; r--) {
            double s = b[r];
            for (int j = r + 1; j < k; j++) s -= a[r * k + j] * b[j];
            b[r] = s / a[r * k + r];
        }
        return true;
    }
    arma::vec parameters;
    bool ok = true;
};
} }
