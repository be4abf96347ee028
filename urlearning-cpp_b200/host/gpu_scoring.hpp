// gpu_scoring.hpp — the reference's plugin/operator interface for the local-score path, backed by liburlgpu.
//
//   scoring::ScoringFunction      scoring_function/scoring_function.h:16-24  (calculateScore per set; BIC, fNML, BDeu, cBIC)
//   scoring::ScoreCalculator      scoring_function/score_calculator.{h,cpp}  (calculateScores per variable, prune)
//   FloatMap                      base/typedefs.h:816                        (per-variable score cache)
//
// Same names, argument meaning and return conventions as the reference: calculateScore returns the score
// (< 0 valid, BIC) or -the_score (cBIC); calculateScores fills the cache with the entries the reference would
// store.  Errors are C++ exceptions (std::runtime_error), as in the reference.  There is no CPU path: every
// call goes through the C ABI in include/urlgpu.h and throws if the device is missing.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "urlearning_host.hpp"
#include "urlgpu.h"

namespace scoring {

using urlhost::Varset;
typedef Varset varset;

// FloatMap stand-in: the entries of one variable in canonical order (|S|, mask) plus a hash index.  The reference's
// boost::unordered_map iteration order is an artefact of Boost's hash; see SURVEY.md Q4.
struct VarsetHash {
    size_t operator()(const varset &k) const {
        uint64_t h = 0x9e3779b97f4a7c15ull;
        for (auto w : k.w) { h ^= w + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); }
        return (size_t)h;
    }
};
struct FloatMap {
    std::vector<varset> keys;
    std::vector<float> values;
    std::unordered_map<varset, size_t, VarsetHash> index;   // built on demand (find / put)
    size_t size() const { return keys.size(); }
    void clear() { keys.clear(); values.clear(); index.clear(); }
    void reindex() { if (index.size() != keys.size()) { index.clear(); for (size_t i = 0; i < keys.size(); i++) index[keys[i]] = i; } }
    const float *find(const varset &k) {
        reindex();
        auto it = index.find(k);
        return it == index.end() ? nullptr : &values[it->second];
    }
    void put(const varset &k, float v) { // cache[k] = v
        reindex();
        auto it = index.find(k);
        if (it != index.end()) values[it->second] = v;
        else { index[k] = keys.size(); keys.push_back(k); values.push_back(v); }
    }
};

class ScoringFunction {
public:
    virtual ~ScoringFunction() {}
    virtual float calculateScore(int variable, varset parents, FloatMap &cache) = 0; // scoring_function.h:19
    virtual int scoreType() const = 0;
    virtual double getLambda() const { return 0; }
    virtual urlgpu_ctx *context() = 0;
};

class GpuContext {
public:
    explicit GpuContext(int device) {
        if (urlgpu_create(&ctx, device) != URLGPU_OK) throw std::runtime_error(std::string("urlgpu: ") + urlgpu_last_error(nullptr));
    }
    ~GpuContext() { urlgpu_destroy(ctx); }
    void check(int rc) const { if (rc != URLGPU_OK) throw std::runtime_error(std::string("urlgpu: ") + urlgpu_last_error(ctx)); }
    urlgpu_ctx *ctx = nullptr;
};

// BICScoringFunction (bic_scoring_function.cpp) on the device
class GpuBICScoringFunction : public ScoringFunction {
public:
    GpuBICScoringFunction(GpuContext &g, const urlhost::BayesianNetwork &network, int recordCount) : g(g) {
        const int p = network.size();
        std::vector<uint8_t> codes((size_t)p * recordCount);
        std::vector<int32_t> card(p);
        for (int i = 0; i < p; i++) {
            card[i] = network.getCardinality(i);
            if (card[i] > 256) throw std::runtime_error("Variable '" + network.get(i).name + "' has more than 256 values");
            for (int r = 0; r < recordCount; r++) codes[(size_t)i * recordCount + r] = (uint8_t)network.codes[i][r];
        }
        g.check(urlgpu_set_discrete(g.ctx, codes.data(), recordCount, p, card.data()));
    }
    // the same from packed codes (column-major, value indices in first-appearance order)
    GpuBICScoringFunction(GpuContext &g, const uint8_t *codes, int64_t recordCount, int p, const int32_t *card) : g(g) {
        g.check(urlgpu_set_discrete(g.ctx, codes, recordCount, p, card));
    }
    // a worker thread on the same device borrows the owner's device copy of the data set (no second upload)
    GpuBICScoringFunction(GpuContext &g, GpuContext &owner) : g(g) { g.check(urlgpu_share_discrete(g.ctx, owner.ctx)); }
    float calculateScore(int variable, varset parents, FloatMap &) override {
        float s;
        g.check(urlgpu_score_one(g.ctx, variable, parents.w, urlhost::kVarsetWords, URLGPU_BIC, 0.0, &s, nullptr));
        return s;
    }
    int scoreType() const override { return URLGPU_BIC; }
    urlgpu_ctx *context() override { return g.ctx; }
private:
    GpuContext &g;
};

// fNMLScoringFunction (fnml_scoring_function.cpp) on the device: the same counts as BIC, the regret tables instead of the penalty
class GpufNMLScoringFunction : public ScoringFunction {
public:
    GpufNMLScoringFunction(GpuContext &g, const uint8_t *codes, int64_t recordCount, int p, const int32_t *card) : g(g) {
        g.check(urlgpu_set_discrete(g.ctx, codes, recordCount, p, card));
    }
    GpufNMLScoringFunction(GpuContext &g, GpuContext &owner) : g(g) { g.check(urlgpu_share_discrete(g.ctx, owner.ctx)); }
    float calculateScore(int variable, varset parents, FloatMap &) override {
        float s;
        g.check(urlgpu_score_one(g.ctx, variable, parents.w, urlhost::kVarsetWords, URLGPU_FNML, 0.0, &s, nullptr));
        return s;
    }
    int scoreType() const override { return URLGPU_FNML; }
    urlgpu_ctx *context() override { return g.ctx; }
private:
    GpuContext &g;
};

// BDeuScoringFunction (bdeu_scoring_function.cpp, deCampos pruning off) on the device; getLambda() carries the equivalent sample size
class GpuBDeuScoringFunction : public ScoringFunction {
public:
    GpuBDeuScoringFunction(GpuContext &g, float ess, const uint8_t *codes, int64_t recordCount, int p, const int32_t *card) : g(g), ess(ess) {
        g.check(urlgpu_set_discrete(g.ctx, codes, recordCount, p, card));
    }
    GpuBDeuScoringFunction(GpuContext &g, float ess, GpuContext &owner) : g(g), ess(ess) { g.check(urlgpu_share_discrete(g.ctx, owner.ctx)); }
    float calculateScore(int variable, varset parents, FloatMap &) override {
        float s;
        g.check(urlgpu_score_one(g.ctx, variable, parents.w, urlhost::kVarsetWords, URLGPU_BDEU, ess, &s, nullptr));
        return s;
    }
    int scoreType() const override { return URLGPU_BDEU; }
    double getLambda() const override { return ess; }
    urlgpu_ctx *context() override { return g.ctx; }
private:
    GpuContext &g;
    float ess;
};

// BIC_OLS_Function (BIC_OLS.cpp) on the device
class GpuBICOLSFunction : public ScoringFunction {
public:
    GpuBICOLSFunction(GpuContext &g, const urlhost::RecordFile &rf, double lambda) : g(g), lambda(lambda) {
        const int n = rf.size(), p = (int)rf.records[0].size();
        std::vector<double> x((size_t)n * p);
        for (int i = 0; i < p; i++)
            for (int r = 0; r < n; r++) x[(size_t)i * n + r] = strtod(rf.records[r][i].c_str(), nullptr); // mlpack::data::Load, BIC_OLS.cpp:48
        g.check(urlgpu_set_continuous(g.ctx, x.data(), n, p));
    }
    GpuBICOLSFunction(GpuContext &g, const double *x_colmajor, int64_t n, int p, double lambda) : g(g), lambda(lambda) {
        g.check(urlgpu_set_continuous(g.ctx, x_colmajor, n, p));
    }
    // a worker thread installs the owner's Gram (the rows are not needed for scoring)
    GpuBICOLSFunction(GpuContext &g, GpuContext &owner, int64_t n, int p, double lambda) : g(g), lambda(lambda) {
        std::vector<double> gram((size_t)p * p);
        owner.check(urlgpu_get_gram(owner.ctx, gram.data()));
        g.check(urlgpu_set_gram(g.ctx, gram.data(), n, p));
    }
    // BIC_OLS.cpp:174-276 per set: the score comes from the device, the acceptance test against the best cached subset
    // (find_best_subset_score :125-172, "clean" recursion) and the callee-side store (:249) run here on the
    // caller's cache, so this plug-in behaves like the reference's under the reference's own enumeration loop
    // (score_calculator.cpp:54-135).  The batched path (ScoreCalculator::calculateScores) does all of this on the device.
    float calculateScore(int variable, varset parents, FloatMap &cache) override {
        float s;
        g.check(urlgpu_score_one(g.ctx, variable, parents.w, urlhost::kVarsetWords, URLGPU_CBIC, lambda, &s, nullptr));
        const float the_score = -s;
        parents.clear(variable);
        const int num_parents = parents.cardinality();
        if (num_parents > 0 && the_score >= 0.0f) return -the_score;                      // :213-224 (bic_threshold = 0, :57)
        std::unordered_map<varset, float, VarsetHash> memo;
        const float best = bestSubsetScore(parents, cache, memo);
        if (num_parents > 0 && best + 0.0f >= -the_score) return -the_score;              // :234-246: dominated, stored nowhere
        cache.put(parents, -the_score);                                                   // :249
        return -the_score;
    }
    int scoreType() const override { return URLGPU_CBIC; }
    double getLambda() const override { return lambda; }
    urlgpu_ctx *context() override { return g.ctx; }
private:
    // F(S) = max(0, max over i in S with S\i non-empty of g(S\i)),  g(T) = cached(T) ? value : F(T)
    static float bestSubsetScore(const varset &parents, FloatMap &cache, std::unordered_map<varset, float, VarsetHash> &memo) {
        float best = 0;
        for (int i = 0; i < urlhost::kVarsetWords * 64; i++) {
            if (!parents.get(i)) continue;
            varset thin = parents;
            thin.clear(i);
            if (thin.cardinality() == 0) continue;                                        // `checked` is seeded with the empty set (:231)
            float sc;
            if (const float *hit = cache.find(thin)) sc = *hit;
            else {
                auto it = memo.find(thin);
                if (it != memo.end()) sc = it->second;
                else { sc = bestSubsetScore(thin, cache, memo); memo[thin] = sc; }
            }
            if (sc > best) best = sc;
        }
        return best;
    }
    GpuContext &g;
    double lambda;
};

class ScoreCalculator {
public:
    ScoreCalculator(ScoringFunction *scoringFunction, int maxParents, int variableCount, bool prune)
        : scoringFunction(scoringFunction), maxParents(maxParents), variableCount(variableCount), pruneFlag(prune) {}

    // score_calculator.cpp:33-135: every subset of neighbors\{variable} with <= maxParents members, stored under
    // the reference's rule; with prune, score_calculator.cpp:150-197 applied on the device as well.
    void calculateScores(int variable, FloatMap &cache, const varset &neighbors) {
        Pending pd = beginScores(variable, neighbors);
        finishScores(pd, cache);
    }

    // The same call split in two so a driver can keep the GPU busy: beginScores enqueues the scoring kernels and the
    // on-device compaction of the cache (no host synchronisation); finishScores waits for that variable only and
    // copies its cache out.  scoringThread calls beginScores(v+1) before finishScores(v).
    struct Pending { urlgpu_result *res = nullptr; int variable = -1; };
    Pending beginScores(int variable, const varset &neighbors) {
        urlgpu_ctx *ctx = scoringFunction->context();
        Pending pd;
        pd.variable = variable;
        unsigned flags = (pruneFlag ? URLGPU_PRUNE_DOMINATED : URLGPU_KEEP_ALL) | extraFlags;
        check(ctx, urlgpu_score_variable(ctx, variable, neighbors.w, urlhost::kVarsetWords, maxParents, scoringFunction->scoreType(),
                                         scoringFunction->getLambda(), flags, &pd.res));
        int rc = urlgpu_result_prefetch(pd.res);
        if (rc != URLGPU_OK) { urlgpu_result_free(pd.res); check(ctx, rc); }
        return pd;
    }
    void finishScores(Pending &pd, FloatMap &cache) {
        urlgpu_ctx *ctx = scoringFunction->context();
        urlgpu_result *res = pd.res;
        pd.res = nullptr;
        uint64_t n = 0;
        int rc = urlgpu_result_count(res, &n);
        if (rc == URLGPU_OK) {
            std::vector<uint64_t> masks(n * urlhost::kVarsetWords);
            cache.keys.resize(n);
            cache.values.resize(n);
            rc = urlgpu_result_fetch(res, 0, n, masks.data(), cache.values.data());
            for (uint64_t i = 0; i < n && rc == URLGPU_OK; i++) memcpy(cache.keys[i].w, &masks[i * urlhost::kVarsetWords], sizeof(cache.keys[i].w));
        }
        urlgpu_result_scored(res, &lastScored);
        urlgpu_result_free(res);
        check(ctx, rc);
    }
    uint64_t lastScored = 0;
    unsigned extraFlags = 0;   // e.g. URLGPU_CBIC_ACCEPT_LITERAL

private:
    static void check(urlgpu_ctx *ctx, int rc) { if (rc != URLGPU_OK) throw std::runtime_error(std::string("urlgpu: ") + urlgpu_last_error(ctx)); }
    ScoringFunction *scoringFunction;
    int maxParents, variableCount;
    bool pruneFlag;
};

} // namespace scoring
