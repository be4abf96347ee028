/*
 * oracle_cli.cpp — command-line front end of the CPU oracle (TEST INFRASTRUCTURE).
 * Mirrors the flags of the reference's `score` (score_main.cpp:216-235) that matter to the hot path,
 * plus oracle-only switches (--prune, --accept, --bic-mode, --from-gram).
 */
#include "oracle.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char **argv) {
    orc_options o;
    memset(&o, 0, sizeof o);
    o.function = "BIC"; o.delimiter = ','; o.lambda = 0.5; o.threads = 1;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char *longname) -> const char * {
            std::string l = std::string("--") + longname + "=";
            if (a.rfind(l, 0) == 0) return argv[i] + l.size();
            if (i + 1 < argc) return argv[++i];
            fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2);
        };
        if (a == "-k" || a.rfind("--skeleton", 0) == 0) o.skeleton = val("skeleton");
        else if (a == "-f" || a.rfind("--function", 0) == 0) o.function = val("function");
        else if (a == "-l" || a.rfind("--lambda", 0) == 0) o.lambda = atof(val("lambda"));
        else if (a == "-p" || a.rfind("--maxParents", 0) == 0) o.max_parents = atoi(val("maxParents"));
        else if (a == "-t" || a.rfind("--threads", 0) == 0) o.threads = atoi(val("threads"));
        else if (a == "-d" || a.rfind("--delimiter", 0) == 0) o.delimiter = val("delimiter")[0];
        else if (a == "-s" || a == "--hasHeader") o.has_header = 1;
        else if (a == "--prune") o.prune = 1;
        else if (a.rfind("--accept", 0) == 0) o.accept_mode = strcmp(val("accept"), "literal-zero") == 0;
        else if (a.rfind("--bic-mode", 0) == 0) o.bic_mode = strcmp(val("bic-mode"), "literal") == 0;
        else if (a == "--from-gram") o.cbic_from_gram = 1;
        else if (a == "-o" || a == "--doNotPrune") {}
        else if (a[0] == '-') { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
        else pos.push_back(a);
    }
    if (pos.size() != 2) { fprintf(stderr, "usage: oracle_score input.csv output.pss [-s] [-k skel] [-f BIC|cBIC] [--lambda=L] [-p K] [-t T] [--prune]\n"); return 2; }
    o.input = pos[0].c_str(); o.output = pos[1].c_str();
    long long n = orc_score_file(&o);
    if (n < 0) { fprintf(stderr, "oracle error: %s\n", orc_last_error()); return 1; }
    printf("oracle: wrote %lld scores to %s\n", n, o.output);
    return 0;
}
