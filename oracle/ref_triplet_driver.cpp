/*
 * ref_triplet_driver.cpp — main() around the REFERENCE's own Triplet A* driver (TEST INFRASTRUCTURE).
 *
 * oracle/ref.mk compiles /root/reference/urlearning/astar/triplet_astar.cpp where it lies (its own main() renamed away with
 * -Dmain=..., Boost replaced by the shims in oracle/shim/) together with the reference's pattern databases, priority queue,
 * score cache and sparse parent structures.  This file only sets the option globals that program_options would have set
 * (triplet_astar.cpp:1624-1687, defaults as there: -b list, -e static, -a 2) and calls the reference's astar().
 * Usage: ref_triplet <scores.pss> <skeleton> <netFile>   ->   <netFile>.csv, element (i, j) = 1 iff i -> j.
 */
#include <cstdio>
#include <string>

extern std::string scoreFile, skeletonFile, netFile, bestScoreCalculator, heuristicType, heuristicArgument, ancestorsArgument, sccArgument;
extern bool outOfTime;
void astar();

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: ref_triplet <scores.pss> <skeleton> <netFile>\n"); return 2; }
    scoreFile = argv[1];
    skeletonFile = argv[2];
    netFile = argv[3];
    bestScoreCalculator = "list";
    heuristicType = "static";
    heuristicArgument = "2";
    ancestorsArgument = "";
    sccArgument = "";
    outOfTime = false;
    astar();
    return 0;
}
