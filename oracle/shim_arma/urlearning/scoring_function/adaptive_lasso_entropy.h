// shim: see lasso_entropy_scoring_function.h in this directory
#pragma once
#include "urlearning/scoring_function/lasso_entropy_scoring_function.h"
namespace scoring {
class AdaptiveLassoEntropyScoringFunction : public ScoringFunction {
public:
    AdaptiveLassoEntropyScoringFunction(datastructures::BayesianNetwork &, int, std::string, double, Constraints *, bool, bool, const datastructures::Skeleton * = NULL) { throw std::runtime_error("the lasso scoring functions are not built in this pin"); }
    float calculateScore(int, varset, FloatMap &) { return 0; }
};
}
