/*
 * ref_cbic_driver.cpp — C entry points around the REFERENCE's own continuous-BIC scoring function (TEST INFRASTRUCTURE).
 *
 * oracle/ref.mk compiles /root/reference/urlearning/scoring_function/BIC_OLS.cpp and score_calculator.cpp where they lie,
 * against oracle/shim (Boost) and oracle/shim_arma (a minimal Armadillo / mlpack: matrices, mean, var, a no-intercept linear
 * regression — their published algorithms in the oracle's arithmetic order).  What this pins is the reference's OWN code:
 * the standardisation (BIC_OLS.cpp:66-80), the score formula (:277-389), the acceptance test with its recursion exactly as
 * written (:125-276, arma::uvec zero-filled as in Armadillo >= 10.5) and the enumeration / store loop that drives it
 * (score_calculator.cpp:54-135) — not Armadillo's or mlpack's floating-point arithmetic, which stays restated.
 */
#include <cstdint>
#include <string>
#include <vector>

#include "urlearning/base/bayesian_network.h"
#include "urlearning/base/record_file.h"
#include "urlearning/scoring_function/BIC_OLS.h"
#include "urlearning/scoring_function/score_calculator.h"

namespace {
struct RefC {
    datastructures::RecordFile *recordFile;
    datastructures::BayesianNetwork *network;
    scoring::BIC_OLS_Function *sf;
};
}

extern "C" {

/* score_main.cpp:283-351 with -f cBIC: RecordFile -> BayesianNetwork (names and count only) -> BIC_OLS_Function(file, lambda) */
void *refc_open(const char *csv, double lambda) {
    RefC *r = new RefC(); /* like the reference, nothing here is ever freed */
    r->recordFile = new datastructures::RecordFile(csv, ',', false);
    r->recordFile->read();
    if (r->recordFile->size() == 0) return NULL;
    r->network = new datastructures::BayesianNetwork();
    r->network->initialize(*r->recordFile);
    r->sf = new scoring::BIC_OLS_Function(*r->network, csv, NULL, false, lambda);
    return r;
}
int refc_p(void *h) { return ((RefC *)h)->network->size(); }

/* calculateScore on an empty cache: -the_score (BIC_OLS.cpp:174-276) */
float refc_calculate_score(void *h, int variable, uint64_t parents) {
    FloatMap cache;
    return ((RefC *)h)->sf->calculateScore(variable, parents, cache);
}

/* scoringThread for one variable (score_main.cpp:143-171): the reference's own enumeration loop, its calls of calculateScore
 * with the growing cache, the callee-side and caller-side stores; optionally the commented-out prune.  Entries come back in
 * the FloatMap's iteration order. */
int64_t refc_score_variable(void *h, int variable, uint64_t neighbor_bits, int maxParents, int prune, uint64_t *masks, float *scores, int64_t cap) {
    RefC *r = (RefC *)h;
    scoring::ScoreCalculator sc(r->sf, maxParents, r->network->size(), -1, NULL);
    FloatMap cache;
    varset nb = neighbor_bits;
    sc.calculateScores(variable, cache, nb);
    if (prune) sc.prune(cache);
    int64_t i = 0;
    for (auto it = cache.begin(); it != cache.end(); ++it, ++i)
        if (i < cap) { masks[i] = it->first; scores[i] = it->second; }
    return (int64_t)cache.size();
}

} /* extern "C" */
