// shim: see lasso_entropy_scoring_function.h in this directory
#pragma once
#include "urlearning/scoring_function/lasso_entropy_scoring_function.h"
namespace scoring {
class AdaptiveLassoEntropyScoringFunction : public ScoringFunction {
public:
    AdaptiveLassoEntropyScoringFunction(datastructures::BayesianNetwork &, int, std::string, double, Constraints *, bool, bool, const datastructures::Skeleton * = NULL) {}
    float calculateScore(int, varset, FloatMap &) { throw std::runtime_error("the lasso scoring functions are not built in this pin"); }
    void post_processing(std::vector<int> &, std::vector<varset> &) {}
};
}
