/*
 * urlsearch.h — C ABI of the host-side consumers of the `.pss` (liburlsearch.so; no CUDA, no GPU needed).
 *
 * the search side of ninalu/urlearning-cpp stays on the host and reads the GPU-written `.pss`
 * unchanged.  This library is the boost-free restatement of exactly the pieces needed to check "the downstream A* DAG is
 * identical" inside this repository (urlearning-cpp_b200/host/search_host.hpp); the `astar` binary is its CLI.
 * What each entry point restates (paths relative to /root/reference/urlearning/):
 *
 *   urlsearch_open          scoring::ScoreCache::read                      score_cache/score_cache.cpp:55-162
 *   urlsearch_entries       FloatMap of one variable (scores negated on read, :151)
 *   urlsearch_best_scores   BestScoreCalculator::getScore(pars)            score_cache/sparse_parent_list.cpp:44-55,
 *                                                                          score_cache/sparse_parent_bitwise.cpp:90-110
 *   urlsearch_astar         astar() / run_astar_on_one_scc                 astar/astar_main.cpp:216-546, 548-645
 *                           with heuristics::StaticPatternDatabase         heuristic/static_pattern_database.cpp:83-248
 *                           and PriorityQueue / CompareNodeStar            priority_queue/priority_queue-inl.h, base/node.h:124-135
 *   urlsearch_triplet       the Triplet A* driver: astar(), process_triple  astar/triplet_astar.cpp:991-1622, 811-989
 * All functions return 0 (or a count) on success and a negative value on error (message via urlsearch_last_error).
 */
#ifndef URLSEARCH_H
#define URLSEARCH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct urlsearch_cache urlsearch_cache;
urlsearch_cache *urlsearch_open(const char *pss_path);            /* NULL on error: urlsearch_last_error(NULL) */
void urlsearch_close(urlsearch_cache *c);
const char *urlsearch_last_error(urlsearch_cache *c);
int urlsearch_variable_count(urlsearch_cache *c);
const char *urlsearch_name(urlsearch_cache *c, int variable);
int urlsearch_arity(urlsearch_cache *c, int variable);
const char *urlsearch_meta(urlsearch_cache *c, const char *key);  /* "" when absent */
/* entries of one variable sorted by (score ascending, |S|, mask): the order of the sparse parent list.  Returns the count. */
int64_t urlsearch_entries(urlsearch_cache *c, int variable, uint64_t *masks, float *scores, int64_t cap);
/* type: "list" or "bitwise".  best[i] = FLT_MAX and parents[i] = 0 when no cached set is a subset of queries[i]. */
int urlsearch_best_scores(urlsearch_cache *c, const char *type, int variable, const uint64_t *queries, int64_t nq, float *best, uint64_t *parents);
/* skeleton_file NULL or "": no skeleton.  parents[v] = optimal parent set; returns the number of components searched. */
int urlsearch_astar(urlsearch_cache *c, const char *type, int pd_count, const char *skeleton_file, float *total_cost, uint64_t *parents, int *nodes_expanded);
/* Triplet A*: directed[i*p + j] = 1 iff i -> j, both directions set for an undirected edge (the reference's <netFile>.csv,
 * README.md:40-45).  A skeleton file is required.  stats (optional, 5 ints): triples searched, colliders found, edges found
 * outside the skeleton, edges oriented by the rules, A* nodes expanded. */
int urlsearch_triplet(urlsearch_cache *c, const char *type, int pd_count, const char *skeleton_file, int32_t *directed, int *stats);
#ifdef __cplusplus
}
#endif
#endif
