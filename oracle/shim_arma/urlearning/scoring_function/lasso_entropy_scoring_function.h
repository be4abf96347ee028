// shim: score_main.cpp can construct the lasso scoring functions (mlpack LARS); they are outside this project's path.  These
// declarations let the reference's main() compile; choosing -f lasso / adaptive throws.  TEST INFRASTRUCTURE.
#pragma once
#include <stdexcept>
#include <string>
#include "urlearning/base/bayesian_network.h"
#include "urlearning/base/skeleton.hpp"
#include "urlearning/scoring_function/constraints.h"
#include "urlearning/scoring_function/scoring_function.h"
namespace scoring {
class LassoEntropyScoringFunction : public ScoringFunction {
public:
    LassoEntropyScoringFunction(datastructures::BayesianNetwork &, int, std::string, double, Constraints *, bool) { throw std::runtime_error("the lasso scoring functions are not built in this pin"); }
    float calculateScore(int, varset, FloatMap &) { return 0; }
};
}
