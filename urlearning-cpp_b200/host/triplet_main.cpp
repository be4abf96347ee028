// triplet_main.cpp — the `triplet_astar` binary: same positional score file and -k / -n / -b / -a options as the reference's
// urlearning/astar/triplet_astar.cpp:1624-1687; writes <netFile>.csv, element (i, j) = 1 iff i -> j, both set for an
// undirected edge (README.md:40-45).  Host code: it consumes the `.pss` the GPU `score` binary writes.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <memory>

#include "triplet_host.hpp"
#include "urlearning_host.hpp"

using namespace urlsearch;

int main(int argc, char **argv) {
    try {
        std::string scoreFile, skeletonFile, netFile, bestScore = "list", heuristicArgument = "2", ignored;
        bool quiet = false;
        for (int i = 1; i < argc; i++) {
            const std::string a = argv[i];
            auto take = [&](const char *s, const char *l, std::string &dst) {
                const std::string lo = std::string("--") + l;
                if (a == s || a == lo) { if (i + 1 >= argc) throw std::runtime_error("missing value for " + a); dst = argv[++i]; return true; }
                if (a.rfind(lo + "=", 0) == 0) { dst = a.substr(lo.size() + 1); return true; }
                return false;
            };
            if (a == "-h" || a == "--help") {
                printf("usage: triplet_astar <scores.pss> -k <skeleton> -n <netFile> [-b list|bitwise] [-a <pattern databases>] [--quiet]\n");
                return 0;
            }
            if (take("-k", "skeleton", skeletonFile) || take("-n", "netFile", netFile) || take("-b", "bestScore", bestScore) || take("-a", "argument", heuristicArgument)) {}
            else if (take("-f", "scoring_function", ignored) || take("-i", "raw_inputFile", ignored) || take("-l", "lambda", ignored) || take("-w", "scoreType", ignored) ||
                     take("-r", "runningTime", ignored) || take("-e", "heuristic", ignored)) {}
            else if (a == "--adaptive") {}
            else if (a == "--quiet") quiet = true;
            else if (a.size() > 1 && a[0] == '-') throw std::runtime_error("unrecognised option '" + a + "'");
            else if (scoreFile.empty()) scoreFile = a;
        }
        if (scoreFile.empty()) throw std::runtime_error("the option '--scoreFile' is required but missing");
        if (skeletonFile.empty()) throw std::runtime_error("Triplet A* needs a skeleton (-k): the reference's driver edits an uninitialised one otherwise");
        for (auto &ch : bestScore) ch = (char)std::tolower((unsigned char)ch);
        const auto t0 = std::chrono::steady_clock::now();
        ScoreCache cache;
        cache.read(scoreFile);
        const int variableCount = cache.getVariableCount();
        std::vector<std::unique_ptr<BestScoreCalculator>> own;
        std::vector<BestScoreCalculator *> spgs;
        for (int i = 0; i < variableCount; i++) {
            if (bestScore == "list") own.emplace_back(new SparseParentList(cache.cache[i]));
            else if (bestScore == "bitwise") own.emplace_back(new SparseParentBitwise(cache.cache[i], variableCount));
            else throw std::runtime_error("Invalid BestScore calculator type: '" + bestScore + "'.  Valid options are 'bitwise' and 'list'.");
            spgs.push_back(own.back().get());
        }
        urlhost::Skeleton sk;
        if (skeletonFile.find(".arc") + 4 == skeletonFile.size()) sk.read_arc_list_file(skeletonFile, variableCount);
        else sk.read_matrix_file(skeletonFile, variableCount);
        std::vector<varset> rows;
        for (int v = 0; v < variableCount; v++) rows.push_back(sk.get_neighbors(v).w[0]);
        TripletDriver driver(variableCount, spgs, rows, std::max(1, atoi(heuristicArgument.c_str())));
        const TripletResult r = driver.run();
        if (!quiet)
            printf("Triplet A* (host restatement): %d variables, %d triples, %d colliders, %d edges outside the skeleton, %d edges oriented by rules in %d rounds, %ld nodes expanded, %.3f s\n",
                   variableCount, r.triplesRun, r.vStructures, r.unfaithfulEdges, r.orientedByRules, r.ruleIterations, r.nodesExpanded,
                   std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        if (!netFile.empty()) {
            std::ofstream out(netFile + ".csv", std::ios_base::trunc);
            if (!out.good()) throw std::runtime_error("Could not open the network file: '" + netFile + ".csv'");
            for (int i = 0; i < variableCount; i++)
                for (int j = 0; j < variableCount; j++) out << r.directed[i][j] << (j + 1 < variableCount ? "," : "\n");
        }
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "triplet_astar: %s\n", e.what());
        return 1;
    }
}
