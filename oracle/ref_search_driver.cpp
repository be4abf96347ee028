/*
 * ref_search_driver.cpp — C entry points around the REFERENCE's own search-side classes (TEST INFRASTRUCTURE).
 *
 * oracle/ref.mk compiles, from the sources where they lie under /root/reference/urlearning/:
 *   score_cache/score_cache.cpp (the .pss reader, :55-162), score_cache/sparse_parent_{list,bitwise,tree}.cpp,
 *   heuristic/static_pattern_database.cpp, priority_queue/priority_queue.cpp, base/bayesian_network.cpp, base/skeleton.cpp
 * against the shim headers in oracle/shim/ into oracle/_ref/libref_search.so.  This file is the only non-reference code
 * in it.  The A* loop itself lives in astar_main.cpp next to main() and the mlpack-dependent post-processing, so it cannot be
 * compiled here; refs_astar restates that loop (astar_main.cpp:216-420 and reconstructSolution :140-166) line by line over
 * the reference's OWN Node, PriorityQueue, BestScoreCalculator, StaticPatternDatabase and Skeleton objects.
 * It pins urlearning-cpp_b200/host/search_host.hpp (tests/test_search.py): reader entries, getScore answers, A* cost / DAG.
 */
#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "urlearning/base/node.h"
#include "urlearning/base/skeleton.hpp"
#include "urlearning/base/typedefs.h"
#include "urlearning/base/bayesian_network.h"
#include "urlearning/base/variable.h"
#include "urlearning/heuristic/static_pattern_database.h"
#include "urlearning/priority_queue/priority_queue.h"
#include "urlearning/score_cache/best_score_calculator.h"
#include "urlearning/score_cache/best_score_creator.h"
#include "urlearning/score_cache/score_cache.h"

namespace {
struct Refs {
    scoring::ScoreCache cache;
    std::string err;
};
}

extern "C" {

void *refs_open(const char *pss) {
    Refs *r = new Refs(); /* like the reference, nothing here is ever freed */
    try { r->cache.read(pss); } catch (const std::exception &e) { return NULL; }
    return r;
}
int refs_variable_count(void *h) { return ((Refs *)h)->cache.getVariableCount(); }
const char *refs_name(void *h, int v) { return ((Refs *)h)->cache.getNetwork()->get(v)->getName().c_str(); }
int refs_arity(void *h, int v) { return ((Refs *)h)->cache.getNetwork()->get(v)->getCardinality(); }
/* every entry of variable v's FloatMap (search-side sign: -score); returns the entry count */
int64_t refs_entries(void *h, int v, uint64_t *masks, float *scores, int64_t cap) {
    FloatMap *m = ((Refs *)h)->cache.getCache(v);
    int64_t n = 0;
    for (auto it = m->begin(); it != m->end(); ++it, ++n)
        if (masks && n < cap) { masks[n] = (uint64_t)it->first; scores[n] = it->second; }
    return n;
}

/* BestScoreCalculator::getScore(pars) of the reference's "list" / "bitwise" / "tree" structures for a batch of queries */
int refs_best_scores(void *h, const char *type, int variable, const uint64_t *queries, int64_t nq, float *out, uint64_t *parent_sets) {
    Refs *r = (Refs *)h;
    try {
        std::vector<bestscorecalculators::BestScoreCalculator *> spgs = bestscorecalculators::create(type, r->cache);
        for (int64_t i = 0; i < nq; i++) {
            varset q = queries[i];
            out[i] = spgs[variable]->getScore(q);
            if (parent_sets) parent_sets[i] = out[i] == std::numeric_limits<float>::max() ? 0 : (uint64_t)spgs[variable]->getParents();
        }
    } catch (const std::exception &e) { r->err = e.what(); return -1; }
    return 0;
}

/* astar() :548-645 + run_astar_on_one_scc :216-420 + reconstructSolution :140-166; no ancestors / scc arguments.
 * skeleton_file may be NULL or "".  parents[v] = optimal parent set of v; returns the number of components or -1. */
int refs_astar(void *h, const char *type, int pdCount, const char *skeleton_file, float *total_cost, uint64_t *parents, int *nodes_expanded) {
    Refs *r = (Refs *)h;
    try {
        scoring::ScoreCache &cache = r->cache;
        const int variableCount = cache.getVariableCount();
        VARSET_NEW(ancestors, variableCount);
        VARSET_NEW(scc, variableCount);
        VARSET_SET_ALL(scc, variableCount);
        std::vector<bestscorecalculators::BestScoreCalculator *> spgs = bestscorecalculators::create(type, cache);
        heuristics::StaticPatternDatabase *heuristic = new heuristics::StaticPatternDatabase(spgs.size(), pdCount, false, ancestors, scc);
        heuristic->initialize(spgs);
        datastructures::Skeleton skeleton;
        std::string sf = skeleton_file ? skeleton_file : "";
        if (sf.find(".arc") + 4 == sf.size()) skeleton.read_arc_list_file(sf, variableCount);
        else skeleton.read_matrix_file(sf);
        if (!skeleton.good()) skeleton.set_variable_count(variableCount);
        const std::vector<varset> &scc_list = skeleton.get_scc();
        *total_cost = 0;
        *nodes_expanded = 0;
        for (int v = 0; v < variableCount; v++) parents[v] = 0;
        for (size_t ci = 0; ci < scc_list.size(); ci++) {
            const varset the_scc = scc_list[ci];
            NodeMap generatedNodes;
            init_map(generatedNodes);
            PriorityQueue openList;
            int first_set_bit = VARSET_FIND_NEXT_SET(the_scc, 0);
            byte leaf0(first_set_bit);
            Node *root = new Node(0.0f, 0.0f, ancestors, leaf0);
            openList.push(root);
            Node *goal = NULL;
            VARSET_NEW(allVariables, variableCount);
            VARSET_SET_VALUE(allVariables, ancestors);
            allVariables = VARSET_OR(allVariables, the_scc);
            float upperBound = std::numeric_limits<float>::max();
            bool complete = false;
            VARSET_NEW(zero_varset, variableCount);
            while (openList.size() > 0) {
                Node *u = openList.pop();
                (*nodes_expanded)++;
                varset variables = u->getSubnetwork();
                if (variables == allVariables) { goal = u; break; }
                if (u->getF() > upperBound) break;
                u->setPqPos(-2);
                for (byte leaf = 0; leaf < variableCount; leaf++) {
                    if (VARSET_GET(variables, leaf)) continue;
                    if (!VARSET_GET(the_scc, leaf)) continue;
                    if (skeleton.good() and not VARSET_EQUAL(variables, zero_varset)) {
                        const varset &neighbors = skeleton.get_neighbors(leaf);
                        if (VARSET_EQUAL(VARSET_AND(variables, neighbors), zero_varset)) continue;
                    }
                    VARSET_COPY(variables, newVariables);
                    VARSET_SET(newVariables, leaf);
                    Node *succ = generatedNodes[newVariables];
                    if (succ == NULL) {
                        float leaf_score = spgs[leaf]->getScore(newVariables);
                        float g = u->getG() + leaf_score;
                        complete = false;
                        float hh = heuristic->h(newVariables, complete);
                        succ = new Node(g, hh, newVariables, leaf);
                        openList.push(succ);
                        generatedNodes[newVariables] = succ;
                        continue;
                    }
                    if (succ->getPqPos() == -2) continue;
                    float g = u->getG() + spgs[leaf]->getScore(variables);
                    if (g < succ->getG()) {
                        succ->setLeaf(leaf);
                        succ->setG(g);
                        openList.update(succ);
                    }
                }
            }
            if (goal == NULL) { r->err = "No solution found."; return -1; }
            *total_cost += goal->getG();
            VARSET_COPY(goal->getSubnetwork(), remainingVariables);
            Node *current = goal;
            int count = cardinality(the_scc);
            for (int i = 0; i < count; i++) {
                int leaf = current->getLeaf();
                spgs[leaf]->getScore(remainingVariables);
                parents[leaf] = (uint64_t)spgs[leaf]->getParents();
                VARSET_CLEAR(remainingVariables, leaf);
                current = generatedNodes[remainingVariables];
            }
        }
        return (int)scc_list.size();
    } catch (const std::exception &e) { r->err = e.what(); return -1; }
}

const char *refs_last_error(void *h) { return ((Refs *)h)->err.c_str(); }

} /* extern "C" */
