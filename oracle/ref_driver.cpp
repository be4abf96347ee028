/*
 * ref_driver.cpp — C entry points around the REFERENCE's own classes (TEST INFRASTRUCTURE).
 *
 * oracle/ref.mk compiles the reference's discrete-BIC path from its sources where they lie
 * (/root/reference/urlearning/{base/bayesian_network,base/skeleton,ad_tree/*,scoring_function/
 * {log_likelihood_calculator,bic_scoring_function,score_calculator}}.cpp) against the shim headers in
 * oracle/shim/ (Boost is not installed; the shims re-implement the handful of Boost facilities those files
 * use).  This file is the only non-reference code in oracle/_ref/libref_bic.so: it drives the reference
 * classes exactly as score_main.cpp:283-347 (set-up) and :132-171 (scoringThread) do, and hands the
 * resulting FloatMap back.  It pins the oracle: counts, enumeration, store rule, prune and (to float32
 * summation-order noise) scores are compared against the real reference code in tests/test_ref_pin.py.
 *
 * The continuous path (BIC_OLS.cpp) needs mlpack + Armadillo: it is built separately over a minimal shim of the two
 * (ref_cbic_driver.cpp -> libref_cbic.so), which pins the reference's own code around the regression; the arithmetic
 * inside mlpack / Armadillo stays "parity unpinned" beyond the Figure_1/2 golden DAG/MEC files.
 */
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "urlearning/ad_tree/ad_tree.h"
#include "urlearning/ad_tree/contingency_table_node.h"
#include "urlearning/base/bayesian_network.h"
#include "urlearning/base/record_file.h"
#include "urlearning/base/skeleton.hpp"
#include "urlearning/base/variable.h"
#include "urlearning/scoring_function/bdeu_scoring_function.h"
#include "urlearning/scoring_function/bic_scoring_function.h"
#include "urlearning/scoring_function/fnml_scoring_function.h"
#include "urlearning/scoring_function/log_likelihood_calculator.h"
#include "urlearning/scoring_function/score_calculator.h"

namespace {
struct Ref {
    datastructures::RecordFile *recordFile;
    datastructures::BayesianNetwork *network;
    scoring::ADTree *adTree;
    scoring::LogLikelihoodCalculator *llc;
    scoring::BICScoringFunction *sf;
    std::vector<float> ilogi;
    scoring::fNMLScoringFunction *fnml = NULL;                 /* built on first use */
    std::vector<std::vector<float> *> *regret = NULL;
};
void flatten(scoring::ContingencyTableNode *ct, const std::vector<int> &vars, const std::vector<int> &card, size_t depth, int64_t idx, int64_t stride,
             int32_t *out) {
    if (ct == NULL) return;
    if (ct->isLeaf()) { out[idx] = ct->getValue(); return; }
    for (int k = 0; k < card[depth]; k++) flatten(ct->getChild(k), vars, card, depth + 1, idx + stride * k, stride * card[depth], out);
}
} // namespace

extern "C" {

/* score_main.cpp:283-347 with -f BIC */
void *ref_open(const char *csv, char delimiter, int hasHeader, int rMin) {
    Ref *r = new Ref(); /* like the reference, nothing here is ever freed */
    r->recordFile = new datastructures::RecordFile(csv, delimiter, hasHeader != 0);
    r->recordFile->read();
    if (r->recordFile->size() == 0) return NULL;
    r->network = new datastructures::BayesianNetwork();
    r->network->initialize(*r->recordFile);
    r->adTree = new scoring::ADTree(rMin);
    r->adTree->initialize(*r->network, *r->recordFile);
    r->adTree->createTree();
    r->ilogi = scoring::LogLikelihoodCalculator::getLogCache(r->recordFile->size());
    r->llc = new scoring::LogLikelihoodCalculator(r->adTree, *r->network, r->ilogi);
    r->sf = new scoring::BICScoringFunction(*r->network, *r->recordFile, r->llc, NULL, false);
    return r;
}
int ref_p(void *h) { return ((Ref *)h)->network->size(); }
int ref_n(void *h) { return ((Ref *)h)->recordFile->size(); }
int ref_cardinality(void *h, int v) { return ((Ref *)h)->network->getCardinality(v); }
const char *ref_name(void *h, int v) { return ((Ref *)h)->network->get(v)->getName().c_str(); }
/* value index the reference assigned to record r of variable v (Variable::getValueIndex) */
int ref_code(void *h, int v, int rec) {
    Ref *r = (Ref *)h;
    return r->network->get(v)->getValueIndex(r->recordFile->getRecords()[rec].get(v));
}

/* ScoringFunction::calculateScore (bic_scoring_function.cpp:32-76) */
float ref_calculate_score(void *h, int variable, uint64_t parents) {
    FloatMap cache;
    return ((Ref *)h)->sf->calculateScore(variable, parents, cache);
}

/* fNMLScoringFunction::calculateScore (fnml_scoring_function.cpp:28-74), set up as score_main.cpp:336-355 does */
float ref_fnml_score(void *h, int variable, uint64_t parents) {
    Ref *r = (Ref *)h;
    if (r->fnml == NULL) {
        r->regret = scoring::getRegretCache(r->recordFile->size(), r->network->getMaxCardinality());
        r->fnml = new scoring::fNMLScoringFunction(*r->network, r->llc, NULL, r->regret, false);
    }
    FloatMap cache;
    return r->fnml->calculateScore(variable, parents, cache);
}
/* the reference's regret table entry regret->at(arity)->at(N) */
float ref_regret(void *h, int arity, int N) {
    Ref *r = (Ref *)h;
    if (r->regret == NULL) ref_fnml_score(h, 0, 0);
    return r->regret->at(arity)->at(N);
}
/* BDeuScoringFunction::calculateScore (bdeu_scoring_function.cpp:37-123), set up as score_main.cpp:357-360 does */
float ref_bdeu_score(void *h, int variable, uint64_t parents, float ess) {
    Ref *r = (Ref *)h;
    /* never freed, like everything else here: the function object holds a BayesianNetwork copy whose destructor would
     * release the variables it shares with the original */
    scoring::BDeuScoringFunction *sf = new scoring::BDeuScoringFunction(ess, *r->network, r->adTree, NULL, false);
    FloatMap cache;
    return sf->calculateScore(variable, parents, cache);
}

/* ADTree::makeContab (ad_tree.cpp:95-137) flattened: mixed radix over the set's variables in ascending index,
 * lowest index least significant.  Returns the number of cells. */
int64_t ref_contab(void *h, uint64_t variables, int32_t *out, int64_t cap) {
    Ref *r = (Ref *)h;
    std::vector<int> vars, card;
    int64_t cells = 1;
    for (int i = 0; i < r->network->size(); i++)
        if (variables & (1ULL << i)) { vars.push_back(i); card.push_back(r->network->getCardinality(i)); cells *= card.back(); }
    if (out == NULL) return cells;
    if (cells > cap) return -1;
    memset(out, 0, cells * sizeof(int32_t));
    scoring::ContingencyTableNode *ct = r->adTree->makeContab(variables);
    flatten(ct, vars, card, 0, 0, 1, out);
    delete ct;
    return cells;
}

/* scoringThread for one variable (score_main.cpp:143-171): calculateScores, then (optionally) the commented-out
 * prune.  Entries are returned in the FloatMap's iteration order. */
int64_t ref_score_variable(void *h, int variable, uint64_t neighbor_bits, int maxParents, int prune, uint64_t *masks, float *scores, int64_t cap) {
    Ref *r = (Ref *)h;
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

/* whole-file throughput with the reference's own threading (score_main.cpp:136-139,372-380): returns sets scored */
int64_t ref_score_all(void *h, const uint64_t *neighbor_bits, int maxParents, int threadCount) {
    Ref *r = (Ref *)h;
    const int p = r->network->size();
    std::vector<int64_t> done(threadCount, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < threadCount; t++)
        th.emplace_back([&, t]() {
            scoring::ScoreCalculator sc(r->sf, maxParents, p, -1, NULL);
            for (int variable = 0; variable < p; variable++) {
                if (variable % threadCount != t) continue;
                FloatMap cache;
                varset nb = neighbor_bits[variable];
                sc.calculateScores(variable, cache, nb);
                done[t] += (int64_t)cache.size();
            }
        });
    for (auto &x : th) x.join();
    int64_t total = 0;
    for (auto d : done) total += d;
    return total;
}

/* Skeleton::read_matrix_file / read_arc_list_file + get_neighbors (skeleton.cpp:19-105, skeleton.hpp:57-60) */
int ref_read_skeleton(const char *path, int p, uint64_t *edges_out) {
    datastructures::Skeleton *sk = new datastructures::Skeleton();
    std::string fn(path);
    bool ok;
    if (fn.find(".arc") + 4 == fn.size()) ok = sk->read_arc_list_file(fn, p);
    else ok = sk->read_matrix_file(fn);
    if (!ok) return -1;
    for (int i = 0; i < p; i++) edges_out[i] = sk->get_neighbors(i);
    return 0;
}
}
