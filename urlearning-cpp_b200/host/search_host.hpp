// search_host.hpp — the consumers of the `.pss` on the host, restated without Boost:
//
//   ScoreCache::read            score_cache/score_cache.cpp:55-162   (.pss reader: case-insensitive "var " / "meta" tests,
//                                                                     order-insensitive, scores negated on read :151)
//   SparseParentList            score_cache/sparse_parent_list.cpp:20-55   (entries sorted by score, first subset wins)
//   SparseParentBitwise         score_cache/sparse_parent_bitwise.cpp:24-110 (per parent a bitset over the sorted entries)
//   StaticPatternDatabase       heuristic/static_pattern_database.cpp:52-248 (reverse BFS over each pattern's lattice)
//   PriorityQueue / Node        priority_queue/priority_queue{.cpp,-inl.h}, base/node.h:24-135 (binary heap with positions,
//                                                                     CompareNodeStar: f, then the deeper layer first)
//   run_astar_on_one_scc        astar/astar_main.cpp:216-546         (A* over the order graph, skeleton-restricted leaves)
//
// These stay on the host (BASELINE.json north_star): they consume the GPU-written `.pss` unchanged.  The restatement
// exists so that "the downstream A* DAG is identical" can be checked inside this repository; tests/test_search.py pins
// it against the reference's own classes compiled from their sources (test infrastructure, outside this package).  Variable sets are 64-bit, as in
// the reference (typedefs.h:469).  Equal scores: the reference's entry order among ties is boost::unordered_map
// iteration order followed by an unstable std::sort; here ties are ordered by (|S|, mask), deterministic.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace urlsearch {

typedef uint64_t varset;
typedef std::unordered_map<varset, float> FloatMap;

inline int cardinality(varset v) { return __builtin_popcountll(v); }

// ------------------------------------------------------------------------------------------------ .pss reader
struct ScoreCache {
    std::vector<std::string> names;
    std::vector<int> arity;
    std::map<std::string, int> nameToIndex;        // operator[] on a missing name yields 0, as bayesian_network.cpp:55-57 does
    std::vector<FloatMap> cache;
    std::map<std::string, std::string> meta;
    int getVariableCount() const { return (int)names.size(); }

    static std::string lower(std::string s) { for (auto &c : s) c = (char)std::tolower((unsigned char)c); return s; }
    static bool contains(const std::string &line, const std::string &str) { return lower(line).find(lower(str)) != std::string::npos; } // :46-53
    static std::string trim(const std::string &s) {
        size_t a = 0, b = s.size();
        while (a < b && std::isspace((unsigned char)s[a])) a++;
        while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
        return s.substr(a, b - a);
    }
    // parse(): substr(start), trim, split on any of `delims` with token compression (:30-36)
    static std::vector<std::string> parse(const std::string &line, size_t start, const std::string &delims) {
        const std::string t = trim(start <= line.size() ? line.substr(start) : std::string());
        std::vector<std::string> out(1);
        for (size_t i = 0; i < t.size();) {
            if (delims.find(t[i]) != std::string::npos) {
                out.emplace_back();
                while (i < t.size() && delims.find(t[i]) != std::string::npos) i++;
            } else out.back().push_back(t[i++]);
        }
        return out;
    }
    int addVariable(const std::string &name) {
        if (nameToIndex.count(name)) throw std::runtime_error("Duplicate variable name: '" + name + "'.");
        nameToIndex[name] = (int)names.size();
        names.push_back(name);
        arity.push_back(0);
        return (int)names.size() - 1;
    }

    void read(const std::string &filename) {
        std::ifstream in(filename);
        if (!in.is_open()) throw std::runtime_error("Could not open the score cache file: '" + filename + "'");
        std::string line;
        std::vector<std::string> tokens;
        bool found = false;
        while (std::getline(in, line)) { // META information until the first variable (:72-98)
            if (line.empty() || line[0] == '#') continue;
            if (contains(line, "var ")) { found = true; break; }
            if (!contains(line, "meta")) throw std::runtime_error("Error while parsing META information of network.  Expected META line or Variable.  Line: '" + line + "'");
            tokens = parse(line, 4, "=");
            if (tokens.size() != 2) throw std::runtime_error("Error while parsing META information of network.  Too many tokens.  Line: '" + line + "'");
            meta[trim(tokens[0])] = trim(tokens[1]);
        }
        if (!found) throw std::runtime_error("The score cache file has no variables: '" + filename + "'");
        tokens = parse(line, 0, " ");
        int v = addVariable(tokens.size() > 1 ? tokens[1] : "");
        while (std::getline(in, line)) { // variable names and arities (:105-133)
            if (line.empty() || line[0] == '#') continue;
            if (contains(line, "meta")) {
                tokens = parse(line, 4, "=");
                if (tokens.size() > 1 && contains(tokens[0], "arity")) arity[v] = atoi(tokens[1].c_str());
            }
            if (contains(line, "var ")) {
                tokens = parse(line, 0, " ");
                v = addVariable(tokens.size() > 1 ? tokens[1] : "");
            }
        }
        in.close();
        if (names.size() > 64) throw std::runtime_error("more than 64 variables: the search side keeps the reference's 64-bit variable sets");
        cache.assign(names.size(), FloatMap());
        in.open(filename); // the parent sets (:139-161)
        while (std::getline(in, line)) {
            if (line.empty() || line[0] == '#' || contains(line, "meta")) continue;
            tokens = parse(line, 0, " ");
            if (contains(line, "var ")) { v = nameToIndex[tokens.size() > 1 ? tokens[1] : ""]; continue; }
            varset parents = 0;
            const float score = (float)(-1 * atof(tokens[0].c_str())); // multiply by -1 to minimise (:151)
            for (size_t i = 1; i < tokens.size(); i++) parents |= (varset)1 << nameToIndex[tokens[i]];
            cache[v][parents] = score;
        }
    }
};

// ------------------------------------------------------------------------------------------------ best-score calculators
struct SortedEntries { // the common part of SparseParentList / SparseParentBitwise: entries sorted by score
    std::vector<varset> parents;
    std::vector<float> scores;
    void build(const FloatMap &m) {
        std::vector<std::pair<varset, float>> spg(m.begin(), m.end());
        std::sort(spg.begin(), spg.end(), [](const std::pair<varset, float> &a, const std::pair<varset, float> &b) {
            if (a.second != b.second) return a.second < b.second;   // sparse_parent_list.cpp:7-9
            const int ca = cardinality(a.first), cb = cardinality(b.first);
            return ca != cb ? ca < cb : a.first < b.first;          // deterministic order among equal scores
        });
        for (auto &e : spg) { parents.push_back(e.first); scores.push_back(e.second); }
    }
};

class BestScoreCalculator {
public:
    virtual ~BestScoreCalculator() {}
    virtual float getScore(varset pars) = 0;      // best score among the cached subsets of pars; remembers its index
    virtual varset getParents() const = 0;        // the parent set of the last getScore
    virtual int size() const = 0;
};

class SparseParentList : public BestScoreCalculator {
public:
    explicit SparseParentList(const FloatMap &m) { e.build(m); }
    float getScore(varset pars) override { // sparse_parent_list.cpp:44-55
        for (bestIndex = 0; bestIndex < (int)e.scores.size(); bestIndex++)
            if ((pars & e.parents[bestIndex]) == e.parents[bestIndex]) break;
        if (bestIndex == (int)e.scores.size()) return std::numeric_limits<float>::max();
        return e.scores[bestIndex];
    }
    varset getParents() const override { return bestIndex >= 0 && bestIndex < (int)e.parents.size() ? e.parents[bestIndex] : 0; }
    int size() const override { return (int)e.parents.size(); }
    SortedEntries e;
private:
    int bestIndex = -1;
};

class SparseParentBitwise : public BestScoreCalculator {
public:
    SparseParentBitwise(const FloatMap &m, int variableCount) : variableCount(variableCount) { // sparse_parent_bitwise.cpp:24-88
        e.build(m);
        words = (e.parents.size() + 63) / 64;
        notUsed.assign((size_t)variableCount * words, ~(uint64_t)0);
        for (size_t i = 0; i < e.parents.size(); i++)
            for (int p = 0; p < variableCount; p++)
                if ((e.parents[i] >> p) & 1) notUsed[(size_t)p * words + (i >> 6)] &= ~((uint64_t)1 << (i & 63));
    }
    float getScore(varset pars) override { // :90-110: AND the complements of every parent NOT allowed, first set bit
        bestIndex = -1;
        for (size_t w = 0; w < words; w++) {
            uint64_t x = ~(uint64_t)0;
            for (int p = 0; p < variableCount; p++)
                if (!((pars >> p) & 1)) x &= notUsed[(size_t)p * words + w];
            if (w == words - 1 && (e.parents.size() & 63)) x &= ((uint64_t)1 << (e.parents.size() & 63)) - 1;
            if (x) { bestIndex = (int)(w * 64 + __builtin_ctzll(x)); break; }
        }
        if (bestIndex < 0) return std::numeric_limits<float>::max();
        return e.scores[bestIndex];
    }
    varset getParents() const override { return bestIndex >= 0 ? e.parents[bestIndex] : 0; }
    int size() const override { return (int)e.parents.size(); }
    SortedEntries e;
private:
    int variableCount, bestIndex = -1;
    size_t words = 0;
    std::vector<uint64_t> notUsed;
};

// ------------------------------------------------------------------------------------------------ static pattern database
class StaticPatternDatabase {
public:
    StaticPatternDatabase(int variableCount, int pdCount, varset ancestors, varset scc)
        : variableCount(variableCount), patternDatabaseCount(pdCount), ancestors(ancestors), scc(scc) {}

    void initialize(std::vector<BestScoreCalculator *> &spgs) { // static_pattern_database.cpp:83-137 (isRandom == false)
        int x = 0;
        const varset allVariables = scc;
        const int remainingCount = cardinality(scc);
        int var = scc ? __builtin_ctzll(scc) : -1;
        const int patternDatabaseSize = (int)std::ceil(static_cast<float>(remainingCount) / patternDatabaseCount);
        for (int pd_i = 0; pd_i < patternDatabaseCount; ++pd_i) {
            variableSets.push_back(0);
            int variableSetSize;
            for (variableSetSize = 0; variableSetSize < patternDatabaseSize && x < remainingCount; variableSetSize++) {
                variableSets[pd_i] |= (varset)1 << var;
                varset rest = var + 1 < 64 ? (scc >> (var + 1)) : 0;         // VARSET_FIND_NEXT_SET(scc, var)
                var = rest ? var + 1 + __builtin_ctzll(rest) : -1;
                ++x;
            }
            patternDatabases.emplace_back();
            createPatternDatabase(allVariables, variableSets[pd_i], variableSetSize, spgs, patternDatabases[pd_i]);
        }
    }

    float h(varset variables, bool &complete) const { // :147-176
        float hval = 0;
        const varset mask = variableCount >= 64 ? ~(varset)0 : (((varset)1 << variableCount) - 1);
        const varset remaining = ~variables & mask;
        for (int pd_i = 0; pd_i < patternDatabaseCount; pd_i++) {
            const varset vs = variableSets[pd_i] & remaining;
            auto it = patternDatabases[pd_i].find(vs);
            if (it == patternDatabases[pd_i].end()) return std::numeric_limits<float>::max() / 64.0f;
            if (vs == remaining) { complete = true; return it->second; }
            hval += it->second;
        }
        return hval;
    }
    int size() const { int s = 0; for (auto &pd : patternDatabases) s += (int)pd.size(); return s; }

private:
    void createPatternDatabase(varset allVariables, varset variableSet, int variableSetSize, std::vector<BestScoreCalculator *> &spgs, FloatMap &patternDatabase) { // :178-222
        FloatMap previousLayer;
        previousLayer[allVariables] = 0;
        for (int layer = 0; layer <= variableSetSize; layer++) { // reverse breadth-first search
            FloatMap currentLayer;
            for (auto &kv : previousLayer) {
                expand(kv.first, kv.second, variableSet, spgs, currentLayer);
                varset pattern = variableSet & ~kv.first;
                pattern &= ~ancestors;
                patternDatabase[pattern] = kv.second;
            }
            previousLayer.swap(currentLayer);
        }
        for (auto &kv : previousLayer) patternDatabase[variableSet & ~kv.first] = kv.second;
    }
    void expand(varset subnetwork, float g, varset variableSet, std::vector<BestScoreCalculator *> &spgs, FloatMap &currentLayer) { // :224-248
        for (int leaf = 0; leaf < variableCount; leaf++) {
            if (!((subnetwork >> leaf) & 1) || !((variableSet >> leaf) & 1)) continue;
            const varset parentChoices = subnetwork | ancestors;   // the leaf's own bit is still set: no cached set contains it
            const float newG = spgs[leaf]->getScore(parentChoices) + g;
            const varset next = subnetwork & ~((varset)1 << leaf);
            float &oldG = currentLayer[next];                      // 0 doubles as "absent" (:243-246)
            if (oldG == 0 || newG < oldG) oldG = newG;
        }
    }
    int variableCount, patternDatabaseCount;
    varset ancestors, scc;
    std::vector<varset> variableSets;
    std::vector<FloatMap> patternDatabases;
};

// ------------------------------------------------------------------------------------------------ node + priority queue
struct Node { // base/node.h:24-118
    float g, h;
    varset subnetwork;
    uint8_t leaf;
    int pqPos;
    Node(float g, float h, varset s, uint8_t leaf) : g(g), h(h), subnetwork(s), leaf(leaf), pqPos(0) {}
    float getF() const { return g + h; }
    int getLayer() const { return cardinality(subnetwork) & 0xff; }
};
struct CompareNodeStar { // node.h:124-135
    bool operator()(const Node *a, const Node *b) const {
        const float diff = a->getF() - b->getF();
        if (std::fabs(diff) < std::numeric_limits<float>::epsilon()) return (b->getLayer() - a->getLayer()) > 0;
        return diff > 0;
    }
};

class PriorityQueue { // priority_queue.cpp:33-64 over the heap routines of priority_queue-inl.h (libstdc++'s, tracking positions)
public:
    int size() const { return (int)pq.size(); }
    void push(Node *n) { pq.push_back(n); pushHeap((long)pq.size() - 1, 0, n); }
    Node *pop() {
        Node *ret = pq.front();
        Node *value = pq.back();               // __pop_heap: the last element goes down from the root
        pq.back() = pq.front();
        adjustHeap(0, (long)pq.size() - 1, value);
        pq.pop_back();
        return ret;
    }
    void update(Node *n) { // the node got a better f: __update_heap (:204-213)
        const long index = n->pqPos, parent = (index - 1) / 2;
        if (index > 0 && cmp(pq[parent], pq[index])) upHeap(index, pq[index]);
        else downHeap(index, pq[index]);
    }
private:
    void pushHeap(long hole, long top, Node *value) { // __push_heap (:17-33)
        long parent = (hole - 1) / 2;
        while (hole > top && cmp(pq[parent], value)) {
            pq[hole] = pq[parent];
            pq[parent]->pqPos = (int)hole;
            hole = parent;
            parent = (hole - 1) / 2;
        }
        pq[hole] = value;
        value->pqPos = (int)hole;
    }
    void adjustHeap(long hole, long len, Node *value) { // __adjust_heap (:68-95)
        const long top = hole;
        long second = hole;
        while (second < (len - 1) / 2) {
            second = 2 * (second + 1);
            if (cmp(pq[second], pq[second - 1])) second--;
            pq[hole] = pq[second];
            pq[second]->pqPos = (int)hole;
            hole = second;
        }
        if ((len & 1) == 0 && second == (len - 2) / 2) {
            second = 2 * (second + 1);
            pq[hole] = pq[second - 1];
            pq[second - 1]->pqPos = (int)hole;
            hole = second - 1;
        }
        pushHeap(hole, top, value);
    }
    void upHeap(long pos, Node *value) { // __up_heap (:147-166)
        long parent = (pos - 1) / 2, index = pos;
        while (index > 0 && cmp(pq[parent], value)) {
            pq[index] = pq[parent];
            pq[parent]->pqPos = (int)index;
            index = parent;
            parent = (parent - 1) / 2;
        }
        if (pos != index) { pq[index] = value; value->pqPos = (int)index; }
    }
    void downHeap(long pos, Node *value) { // __down_heap (:168-196), including its unusual child selection
        const long len = (long)pq.size();
        long index = pos, left = index * 2 + 1, right = index * 2 + 2, largest = len;
        while (index < len) {
            if (right >= len || (left < len && cmp(pq[right], pq[left]))) largest = left;
            if (largest < len && cmp(value, pq[largest])) {
                pq[index] = pq[largest];
                pq[largest]->pqPos = (int)index;
                index = largest;
                left = index * 2 + 1;
                right = index * 2 + 2;
            } else break;
        }
        if (pos != index) pq[index] = value;
    }
    std::vector<Node *> pq;
    CompareNodeStar cmp;
};

// ------------------------------------------------------------------------------------------------ A*
struct AstarResult {
    bool found = false;
    float cost = 0;                       // goal->getG()
    int nodesExpanded = 0;
    std::vector<int> order;               // total ordering of the component (first = a root of the DAG)
    std::vector<varset> parents;          // parents[variable]
};

// astar_main.cpp:216-546 for one connected component `the_scc` of the skeleton (edges empty: no skeleton restriction)
// reopenClosed: the Triplet driver's copy of this loop (astar/triplet_astar.cpp:283-674) lacks the "closed nodes stay closed"
// line and pushes a closed node again when it finds a better g for it (:541-555)
inline AstarResult run_astar_on_one_scc(int variableCount, std::vector<BestScoreCalculator *> &spgs, const StaticPatternDatabase &heuristic, varset ancestors,
                                        varset the_scc, const std::vector<varset> &edges, bool reopenClosed = false) {
    AstarResult out;
    std::unordered_map<varset, Node *> generatedNodes;
    PriorityQueue openList;
    const uint8_t firstLeaf = the_scc ? (uint8_t)__builtin_ctzll(the_scc) : 0;
    Node *root = new Node(0.0f, 0.0f, ancestors, firstLeaf); // not in generatedNodes, as in the reference (:232-233)
    openList.push(root);
    Node *goal = nullptr;
    const varset allVariables = ancestors | the_scc;
    const float upperBound = std::numeric_limits<float>::max();
    const bool skeletonGood = !edges.empty();
    while (openList.size() > 0) {
        Node *u = openList.pop();
        out.nodesExpanded++;
        const varset variables = u->subnetwork;
        if (variables == allVariables) { goal = u; break; }
        if (u->getF() > upperBound) break;
        u->pqPos = -2; // closed
        for (int leaf = 0; leaf < variableCount; leaf++) {
            if ((variables >> leaf) & 1) continue;
            if (!((the_scc >> leaf) & 1)) continue;
            if (skeletonGood && variables != 0 && (variables & edges[leaf]) == 0) continue; // :300-307: the leaf must touch the sub-network
            const varset newVariables = variables | ((varset)1 << leaf);
            Node *&slot = generatedNodes[newVariables];
            if (slot == nullptr) {
                const float g = u->g + spgs[leaf]->getScore(newVariables);
                bool complete = false;
                const float h = heuristic.h(newVariables, complete);
                slot = new Node(g, h, newVariables, (uint8_t)leaf);
                openList.push(slot);
                continue;
            }
            Node *succ = slot;
            if (succ->pqPos == -2 && !reopenClosed) continue; // consistent heuristic: closed nodes stay closed
            const float g = u->g + spgs[leaf]->getScore(variables);
            if (g < succ->g) {
                succ->leaf = (uint8_t)leaf;
                succ->g = g;
                if (succ->pqPos == -2) { succ->pqPos = 0; openList.push(succ); }
                else openList.update(succ);
            }
        }
    }
    if (goal) { // reconstructSolution (:140-166)
        out.found = true;
        out.cost = goal->g;
        const int count = cardinality(the_scc);
        out.order.assign(count, -1);
        out.parents.assign(variableCount, 0);
        varset remaining = goal->subnetwork;
        Node *current = goal;
        for (int i = 0; i < count && current; i++) {
            const int leaf = current->leaf;
            out.order[count - 1 - i] = leaf;
            spgs[leaf]->getScore(remaining);
            out.parents[leaf] = spgs[leaf]->getParents();
            remaining &= ~((varset)1 << leaf);
            auto it = generatedNodes.find(remaining);
            current = it != generatedNodes.end() ? it->second : nullptr; // closedList[ancestors] is NULL after the last leaf
        }
    }
    for (auto &kv : generatedNodes) delete kv.second;
    delete root;
    return out;
}

// connected components of the skeleton in the reference's order (skeleton.cpp:187-230); no skeleton: one component
inline std::vector<varset> components(int variableCount, const std::vector<varset> &edges) {
    std::vector<varset> out;
    if (edges.empty()) { out.push_back(variableCount >= 64 ? ~(varset)0 : (((varset)1 << variableCount) - 1)); return out; }
    varset visited = 0;
    for (int v = 0; v < variableCount; v++) {
        if ((visited >> v) & 1) continue;
        varset comp = 0;
        std::vector<int> stack{v};
        // explore_one_scc is a depth-first recursion in index order; the component as a SET does not depend on the order
        while (!stack.empty()) {
            const int u = stack.back();
            stack.pop_back();
            if ((visited >> u) & 1) continue;
            visited |= (varset)1 << u;
            comp |= (varset)1 << u;
            for (int i = variableCount - 1; i >= 0; i--)
                if (((edges[u] >> i) & 1) && !((visited >> i) & 1)) stack.push_back(i);
        }
        out.push_back(comp);
    }
    return out;
}

} // namespace urlsearch
