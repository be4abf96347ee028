// astar_main.cpp — the `astar` binary: learns the optimal network from a `.pss` exactly as the reference's
// urlearning/astar/astar_main.cpp:548-710 does (sparse parent lists or bitwise, static pattern database heuristic,
// A* per connected component of the skeleton), without Boost.  Host only: the search side stays on the CPU
// (BASELINE.json north_star); this restatement exists so the "downstream DAG identical" check of the GPU-written `.pss`
// runs inside this repository.  Differences, documented in DESIGN.md: only the `static` heuristic and the `list` /
// `bitwise` calculators are offered; the print-only Lasso post-processing (needs mlpack, :482-491) is omitted; -r ignored.
//
//   astar <scores.pss> [-k skeleton] [-b list|bitwise] [-e static] [-a pdCount] [-n netFile]
// Output: "Found solution: <cost>" per component; with -n, <netFile>.csv holds the p x p matrix, (i,j) = 1 iff j is a
// parent of i (astar_main.cpp:514-531; Figure_1/README.txt:1-8).
#include <chrono>
#include <cstring>
#include <iostream>
#include <memory>

#include "search_host.hpp"
#include "urlearning_host.hpp"

using namespace urlsearch;

int main(int argc, char **argv) {
    std::string scoreFile, skeletonFile, bestScore = "list", heuristicType = "static", heuristicArgument = "2", netFile;
    bool quiet = false;
    try {
        for (int i = 1; i < argc; i++) {
            std::string a = argv[i];
            auto take = [&](const char *s, const char *l, std::string &dst) {
                if (a == s || a == std::string("--") + l) { if (i + 1 >= argc) throw std::runtime_error(std::string("missing argument of --") + l); dst = argv[++i]; return true; }
                const std::string pre = std::string("--") + l + "=";
                if (a.rfind(pre, 0) == 0) { dst = a.substr(pre.size()); return true; }
                return false;
            };
            std::string ignored;
            if (a == "-h" || a == "--help") { std::cout << "Learn an optimal Bayesian network using A*.  Example usage: " << argv[0] << " iris.pss [-k skeleton] [-b list|bitwise] [-a 2] [-n net]\n"; return 0; }
            else if (take("-k", "skeleton", skeletonFile)) {}
            else if (take("-b", "bestScore", bestScore)) {}
            else if (take("-e", "heuristic", heuristicType)) {}
            else if (take("-a", "argument", heuristicArgument)) {}
            else if (take("-n", "netFile", netFile)) {}
            else if (take("-f", "scoring_function", ignored) || take("-i", "raw_inputFile", ignored) || take("-l", "lambda", ignored) || take("-w", "scoreType", ignored) ||
                     take("-r", "runningTime", ignored)) {}
            else if (a == "--adaptive") {}
            else if (a == "--quiet") quiet = true;
            else if (a.size() > 1 && a[0] == '-') throw std::runtime_error("unrecognised option '" + a + "'");
            else if (scoreFile.empty()) scoreFile = a;
        }
        if (scoreFile.empty()) throw std::runtime_error("the option '--scoreFile' is required but missing");
        for (auto &ch : bestScore) ch = (char)std::tolower((unsigned char)ch);
        for (auto &ch : heuristicType) ch = (char)std::tolower((unsigned char)ch);
        if (heuristicType != "static") throw std::runtime_error("Invalid heuristic type: '" + heuristicType + "'.  This build offers 'static'.");
        const auto t0 = std::chrono::steady_clock::now();
        printf("URLearning, A* (host restatement)\nDataset: '%s'\nNet file: '%s'\nBest score calculator: '%s'\n", scoreFile.c_str(), netFile.c_str(), bestScore.c_str());
        ScoreCache cache;
        cache.read(scoreFile);
        const int variableCount = cache.getVariableCount();
        printf("Variable count is %d\n", variableCount);
        std::vector<std::unique_ptr<BestScoreCalculator>> own;
        std::vector<BestScoreCalculator *> spgs;
        for (int i = 0; i < variableCount; i++) { // best_score_creator.h:28-47
            if (bestScore == "list") own.emplace_back(new SparseParentList(cache.cache[i]));
            else if (bestScore == "bitwise") own.emplace_back(new SparseParentBitwise(cache.cache[i], variableCount));
            else throw std::runtime_error("Invalid BestScore calculator type: '" + bestScore + "'.  Valid options are 'bitwise' and 'list'.");
            spgs.push_back(own.back().get());
        }
        const varset ancestors = 0;
        const varset scc = variableCount >= 64 ? ~(varset)0 : (((varset)1 << variableCount) - 1);
        StaticPatternDatabase heuristic(variableCount, std::max(1, atoi(heuristicArgument.c_str())), ancestors, scc);
        heuristic.initialize(spgs);
        std::vector<varset> edges;
        if (!skeletonFile.empty()) { // astar_main.cpp:620-626; an unreadable file leaves "no skeleton" in the reference, an error here
            urlhost::Skeleton sk;
            if (skeletonFile.find(".arc") + 4 == skeletonFile.size()) sk.read_arc_list_file(skeletonFile, variableCount);
            else sk.read_matrix_file(skeletonFile, variableCount);
            for (int v = 0; v < variableCount; v++) edges.push_back(sk.get_neighbors(v).w[0]);
        }
        const std::vector<varset> scc_list = components(variableCount, edges);
        printf("num of sccs = %d\n", (int)scc_list.size());
        std::vector<varset> parents(variableCount, 0);
        double total = 0;
        bool all = true;
        for (size_t i = 0; i < scc_list.size(); i++) {
            AstarResult r = run_astar_on_one_scc(variableCount, spgs, heuristic, ancestors, scc_list[i], edges);
            if (!quiet) printf("Nodes expanded: %d\n", r.nodesExpanded);
            if (!r.found) { printf("No solution found.\n"); all = false; continue; }
            printf("Found solution: %f, scc # %d\n", r.cost, (int)i);
            total += r.cost;
            for (int v = 0; v < variableCount; v++) if ((scc_list[i] >> v) & 1) parents[v] = r.parents[v];
            if (!quiet) { printf("total ordering:"); for (int v : r.order) printf(" %d", v); printf("\n"); }
        }
        printf("Total score: %f\n", total);
        if (!netFile.empty() && all) {
            std::ofstream out(netFile + ".csv", std::ios_base::trunc);
            if (!out.good()) throw std::runtime_error("Could not open the network file: '" + netFile + ".csv'");
            for (int v = 0; v < variableCount; v++) {
                for (int i = 0; i < variableCount; i++) out << (((parents[v] >> i) & 1) ? 1 : 0) << (i + 1 < variableCount ? "," : "\n");
            }
        }
        printf("%.3f s wall\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return all ? 0 : 2;
    } catch (const std::exception &e) {
        fprintf(stderr, "astar: %s\n", e.what());
        return 1;
    }
}
