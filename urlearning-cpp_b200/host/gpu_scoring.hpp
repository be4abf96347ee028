// gpu_scoring.hpp — the reference's plugin/operator interface for the local-score path, backed by liburlgpu.
//
//   scoring::ScoringFunction      scoring_function/scoring_function.h:16-24  (calculateScore per set)
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
#include <vector>

#include "urlearning_host.hpp"
#include "urlgpu.h"

namespace scoring {

using urlhost::Varset;
typedef Varset varset;

// FloatMap stand-in: the entries of one variable in canonical order (|S|, mask).  The reference's
// boost::unordered_map iteration order is an artefact of Boost's hash; see SURVEY.md Q4.
struct FloatMap {
    std::vector<varset> keys;
    std::vector<float> values;
    size_t size() const { return keys.size(); }
    void clear() { keys.clear(); values.clear(); }
    const float *find(const varset &k) const {
        for (size_t i = 0; i < keys.size(); i++) if (keys[i] == k) return &values[i];
        return nullptr;
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
    float calculateScore(int variable, varset parents, FloatMap &) override {
        float s;
        g.check(urlgpu_score_one(g.ctx, variable, parents.w, urlhost::kVarsetWords, URLGPU_CBIC, lambda, &s, nullptr));
        return s;
    }
    int scoreType() const override { return URLGPU_CBIC; }
    double getLambda() const override { return lambda; }
    urlgpu_ctx *context() override { return g.ctx; }
private:
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
        unsigned flags = pruneFlag ? URLGPU_PRUNE_DOMINATED : URLGPU_KEEP_ALL;
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

private:
    static void check(urlgpu_ctx *ctx, int rc) { if (rc != URLGPU_OK) throw std::runtime_error(std::string("urlgpu: ") + urlgpu_last_error(ctx)); }
    ScoringFunction *scoringFunction;
    int maxParents, variableCount;
    bool pruneFlag;
};

} // namespace scoring
