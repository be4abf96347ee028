// urlearning_host.hpp — host-side data structures of the `score` binary: the reference's L1 layer
//  re-stated without Boost, widened to multi-word variable sets.
//
//   Varset          base/typedefs.h:469,650-697 (uint64_t varset + VARSET_* macros), widened to 256 variables
//   RecordFile      base/record_file.h:39-54, base/record.h:35-39
//   BayesianNetwork base/bayesian_network.cpp:25-42, base/variable.h:43-64
//   Skeleton        base/skeleton.cpp:19-105, base/skeleton.hpp:57-68
#pragma once
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace urlhost {

constexpr int kVarsetWords = 4; // up to 256 variables

struct Varset {
    uint64_t w[kVarsetWords] = {0, 0, 0, 0};
    bool get(int i) const { return (w[i >> 6] >> (i & 63)) & 1; }
    void set(int i) { w[i >> 6] |= (uint64_t)1 << (i & 63); }
    void clear(int i) { w[i >> 6] &= ~((uint64_t)1 << (i & 63)); }
    void setAll(int n) { for (int i = 0; i < n; i++) set(i); }
    Varset operator|(const Varset &o) const { Varset r; for (int i = 0; i < kVarsetWords; i++) r.w[i] = w[i] | o.w[i]; return r; }
    bool operator==(const Varset &o) const { return memcmp(w, o.w, sizeof w) == 0; }
    int cardinality() const { int c = 0; for (auto x : w) c += __builtin_popcountll(x); return c; }
    bool isSubsetOf(const Varset &o) const { for (int i = 0; i < kVarsetWords; i++) if ((w[i] & o.w[i]) != w[i]) return false; return true; }
    // numeric order of the mask as one big integer (the canonical .pss line order within a layer)
    bool lessThan(const Varset &o) const { for (int i = kVarsetWords - 1; i >= 0; i--) if (w[i] != o.w[i]) return w[i] < o.w[i]; return false; }
};

inline std::string trim(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}

// boost::split(..., is_any_of(delim), token_compress_on): runs of the delimiter count once
inline std::vector<std::string> splitCompress(const std::string &s, char delim) {
    std::vector<std::string> out(1);
    for (size_t i = 0; i < s.size();) {
        if (s[i] == delim) {
            out.emplace_back();
            while (i < s.size() && s[i] == delim) i++;
        } else out.back().push_back(s[i++]);
    }
    return out;
}

class RecordFile {
public:
    RecordFile(std::string filename, char delimiter, bool hasHeader) : filename(std::move(filename)), delimiter(delimiter), hasHeader(hasHeader) {}
    void read() {
        std::ifstream file(filename);
        if (!file.good()) throw std::runtime_error("Could not open the input file: '" + filename + "'");
        std::string line;
        if (hasHeader) { std::getline(file, line); header = splitCompress(trim(line), delimiter); }
        while (std::getline(file, line)) records.push_back(splitCompress(trim(line), delimiter));
        if (records.empty()) throw std::runtime_error("The input file has no records: '" + filename + "'");
        const size_t width = records[0].size();
        for (size_t r = 0; r < records.size(); r++)
            if (records[r].size() < width) throw std::runtime_error("Record " + std::to_string(r + 1) + " has fewer fields than the first record");
    }
    int size() const { return (int)records.size(); }
    bool getHasHeader() const { return hasHeader; }
    std::vector<std::string> header;
    std::vector<std::vector<std::string>> records;
private:
    std::string filename;
    char delimiter;
    bool hasHeader;
};

class Variable {
public:
    std::string name;
    std::vector<std::string> values;                      // first-appearance order (variable.h:43-48)
    std::unordered_map<std::string, int> valueToIndex;
    int getCardinality() const { return (int)values.size(); }
    int addValue(const std::string &v) {
        auto it = valueToIndex.find(v);
        if (it != valueToIndex.end()) return it->second;
        int idx = (int)values.size();
        valueToIndex.emplace(v, idx);
        values.push_back(v);
        return idx;
    }
};

class BayesianNetwork {
public:
    void initialize(const RecordFile &rf) {
        const int p = (int)rf.records[0].size();
        variables.resize(p);
        for (int i = 0; i < p; i++)
            variables[i].name = (rf.getHasHeader() && i < (int)rf.header.size()) ? rf.header[i] : "Variable_" + std::to_string(i);
        codes.assign(p, std::vector<int>(rf.records.size()));
        for (int i = 0; i < p; i++)
            for (size_t r = 0; r < rf.records.size(); r++) codes[i][r] = variables[i].addValue(rf.records[r][i]);
    }
    int size() const { return (int)variables.size(); }
    const Variable &get(int i) const { return variables[i]; }
    int getCardinality(int i) const { return variables[i].getCardinality(); }
    std::vector<Variable> variables;
    std::vector<std::vector<int>> codes; // [variable][record] value index
};

class Skeleton {
public:
    explicit Skeleton(int variableCount = 1) { set_variable_count(variableCount); }
    void set_variable_count(int n) { variableCount = n; all_bit_set = Varset(); all_bit_set.setAll(n); }
    bool good() const { return initialized; }
    const Varset &get_neighbors(int v) const { return initialized ? edges[v] : all_bit_set; }
    // The reference returns false and silently keeps a 1-variable skeleton when the file is unreadable
    // (SURVEY.md Q11); here that is an error.
    void read_matrix_file(const std::string &fn, int p) {
        std::ifstream in(fn);
        if (!in.good()) throw std::runtime_error("Could not open the skeleton file: '" + fn + "'");
        begin(p);
        std::string line;
        int row = 0;
        bool first = true;
        while (std::getline(in, line)) {
            auto tok = tokenize(line, ", \n\r");
            if (first) { first = false; if ((int)tok.size() != p) throw std::runtime_error("Skeleton matrix width differs from the variable count"); }
            int col = 0;
            for (auto &s : tok) {
                // skeleton.cpp:91 `abs(atof(x)) > 0.05` binds ::abs(int) under GCC: x is truncated, so |x| >= 1 or "TRUE"
                if (s == "TRUE" || std::abs((int)atof(s.c_str())) > 0.05) add_edge(row, col);
                col++;
            }
            row++; // blank lines count as rows (skeleton.cpp:84-99)
        }
        initialized = true;
    }
    void read_arc_list_file(const std::string &fn, int p) {
        std::ifstream in(fn);
        if (!in.good()) throw std::runtime_error("Could not open the skeleton file: '" + fn + "'");
        begin(p);
        std::string line;
        while (std::getline(in, line)) {
            auto tok = tokenize(line, ",");
            if (tok.size() < 2) continue;
            int v1 = tok[0].size() > 2 ? atoi(tok[0].c_str() + 2) : 0; // skeleton.cpp:43-44
            int v2 = tok[1].size() > 2 ? atoi(tok[1].c_str() + 2) : 0;
            add_edge(v1 - 1, v2 - 1);
        }
        initialized = true;
    }
private:
    static std::vector<std::string> tokenize(const std::string &s, const char *seps) {
        std::vector<std::string> out;
        std::string cur;
        for (char c : s) {
            if (strchr(seps, c)) { if (!cur.empty()) { out.push_back(cur); cur.clear(); } }
            else cur.push_back(c);
        }
        if (!cur.empty()) out.push_back(cur);
        return out;
    }
    void begin(int p) { set_variable_count(p); edges.assign(p, Varset()); }
    void add_edge(int i, int j) {
        if (i < 0 || j < 0 || i >= variableCount || j >= variableCount) throw std::runtime_error("Skeleton entry out of range");
        edges[i].set(j);
        edges[j].set(i);
    }
    bool initialized = false;
    int variableCount = 1;
    Varset all_bit_set;
    std::vector<Varset> edges;
};

} // namespace urlhost
