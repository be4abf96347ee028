// shim: score_main.cpp can construct the lasso scoring functions (mlpack LARS); they are outside this project's path.  These
// declarations let the reference's main() compile; scoring with -f lasso / adaptive throws; astar's print-only post-processing is a no-op.  TEST INFRASTRUCTURE.
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
    LassoEntropyScoringFunction(datastructures::BayesianNetwork &, int, std::string, double, Constraints *, bool) {}
    float calculateScore(int, varset, FloatMap &) { throw std::runtime_error("the lasso scoring functions are not built in this pin"); }
    void post_processing(std::vector<int> &, std::vector<varset> &) {}   // print-only in the reference (astar_main.cpp:482-491)
};
}
