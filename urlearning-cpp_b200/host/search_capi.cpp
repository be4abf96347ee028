// search_capi.cpp — liburlsearch.so: the C ABI (include/urlsearch.h) of the host-side search restatement in search_host.hpp.
#include <memory>

#include "search_host.hpp"
#include "triplet_host.hpp"
#include "urlearning_host.hpp"
#include "urlsearch.h"

using namespace urlsearch;

struct urlsearch_cache {
    ScoreCache cache;
    std::string err;
};
static thread_local std::string g_open_error;

static std::vector<std::unique_ptr<BestScoreCalculator>> make_spgs(urlsearch_cache *c, const std::string &type) {
    std::vector<std::unique_ptr<BestScoreCalculator>> own;
    const int p = c->cache.getVariableCount();
    for (int i = 0; i < p; i++) {
        if (type == "list") own.emplace_back(new SparseParentList(c->cache.cache[i]));
        else if (type == "bitwise") own.emplace_back(new SparseParentBitwise(c->cache.cache[i], p));
        else throw std::runtime_error("Invalid BestScore calculator type: '" + type + "'.  Valid options are 'bitwise' and 'list'.");
    }
    return own;
}

extern "C" {

urlsearch_cache *urlsearch_open(const char *pss_path) {
    auto *c = new urlsearch_cache();
    try { c->cache.read(pss_path ? pss_path : ""); } catch (const std::exception &e) { g_open_error = e.what(); delete c; return nullptr; }
    return c;
}
void urlsearch_close(urlsearch_cache *c) { delete c; }
const char *urlsearch_last_error(urlsearch_cache *c) { return c ? c->err.c_str() : g_open_error.c_str(); }
int urlsearch_variable_count(urlsearch_cache *c) { return c ? c->cache.getVariableCount() : -1; }
const char *urlsearch_name(urlsearch_cache *c, int v) { return c->cache.names[v].c_str(); }
int urlsearch_arity(urlsearch_cache *c, int v) { return c->cache.arity[v]; }
const char *urlsearch_meta(urlsearch_cache *c, const char *key) {
    auto it = c->cache.meta.find(key);
    return it == c->cache.meta.end() ? "" : it->second.c_str();
}
int64_t urlsearch_entries(urlsearch_cache *c, int v, uint64_t *masks, float *scores, int64_t cap) {
    SortedEntries e;
    e.build(c->cache.cache[v]);
    for (int64_t i = 0; i < (int64_t)e.parents.size() && i < cap && masks; i++) { masks[i] = e.parents[i]; scores[i] = e.scores[i]; }
    return (int64_t)e.parents.size();
}
int urlsearch_best_scores(urlsearch_cache *c, const char *type, int variable, const uint64_t *queries, int64_t nq, float *best, uint64_t *parents) {
    try {
        std::unique_ptr<BestScoreCalculator> spg;
        const std::string t = type ? type : "list";
        if (t == "list") spg.reset(new SparseParentList(c->cache.cache[variable]));
        else if (t == "bitwise") spg.reset(new SparseParentBitwise(c->cache.cache[variable], c->cache.getVariableCount()));
        else throw std::runtime_error("Invalid BestScore calculator type: '" + t + "'.");
        for (int64_t i = 0; i < nq; i++) {
            best[i] = spg->getScore(queries[i]);
            if (parents) parents[i] = best[i] == std::numeric_limits<float>::max() ? 0 : spg->getParents();
        }
    } catch (const std::exception &e) { c->err = e.what(); return -1; }
    return 0;
}
int urlsearch_astar(urlsearch_cache *c, const char *type, int pd_count, const char *skeleton_file, float *total_cost, uint64_t *parents, int *nodes_expanded) {
    try {
        const int p = c->cache.getVariableCount();
        auto own = make_spgs(c, type ? type : "list");
        std::vector<BestScoreCalculator *> spgs;
        for (auto &o : own) spgs.push_back(o.get());
        const varset scc = p >= 64 ? ~(varset)0 : (((varset)1 << p) - 1);
        StaticPatternDatabase heuristic(p, std::max(1, pd_count), 0, scc);
        heuristic.initialize(spgs);
        std::vector<varset> edges;
        if (skeleton_file && *skeleton_file) {
            urlhost::Skeleton sk;
            const std::string sf = skeleton_file;
            if (sf.find(".arc") + 4 == sf.size()) sk.read_arc_list_file(sf, p);
            else sk.read_matrix_file(sf, p);
            for (int v = 0; v < p; v++) edges.push_back(sk.get_neighbors(v).w[0]);
        }
        const std::vector<varset> comps = components(p, edges);
        *total_cost = 0;
        if (nodes_expanded) *nodes_expanded = 0;
        for (int v = 0; v < p; v++) parents[v] = 0;
        for (auto comp : comps) {
            AstarResult r = run_astar_on_one_scc(p, spgs, heuristic, 0, comp, edges);
            if (!r.found) throw std::runtime_error("No solution found.");
            *total_cost += r.cost;
            if (nodes_expanded) *nodes_expanded += r.nodesExpanded;
            for (int v = 0; v < p; v++) if ((comp >> v) & 1) parents[v] = r.parents[v];
        }
        return (int)comps.size();
    } catch (const std::exception &e) { c->err = e.what(); return -1; }
}

// astar/triplet_astar.cpp:991-1622 — the Triplet A* driver (host/triplet_host.hpp)
int urlsearch_triplet(urlsearch_cache *c, const char *type, int pd_count, const char *skeleton_file, int32_t *directed /*[p*p], row-major: [i*p+j] = 1 iff i -> j*/,
                      int *stats /*optional [5]: triples, colliders, edges outside the skeleton, edges oriented by rules, nodes expanded*/) {
    try {
        const int p = c->cache.getVariableCount();
        if (!skeleton_file || !*skeleton_file) throw std::runtime_error("Triplet A* needs a skeleton file");
        auto own = make_spgs(c, type ? type : "list");
        std::vector<BestScoreCalculator *> spgs;
        for (auto &o : own) spgs.push_back(o.get());
        urlhost::Skeleton sk;
        const std::string sf = skeleton_file;
        if (sf.find(".arc") + 4 == sf.size()) sk.read_arc_list_file(sf, p);
        else sk.read_matrix_file(sf, p);
        std::vector<varset> rows;
        for (int v = 0; v < p; v++) rows.push_back(sk.get_neighbors(v).w[0]);
        TripletDriver driver(p, spgs, rows, std::max(1, pd_count));
        const TripletResult r = driver.run();
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) directed[i * p + j] = r.directed[i][j];
        if (stats) { stats[0] = r.triplesRun; stats[1] = r.vStructures; stats[2] = r.unfaithfulEdges; stats[3] = r.orientedByRules; stats[4] = (int)r.nodesExpanded; }
        return 0;
    } catch (const std::exception &e) { c->err = e.what(); return -1; }
}

} // extern "C"
