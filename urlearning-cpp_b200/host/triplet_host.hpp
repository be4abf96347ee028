// triplet_host.hpp — boost-free host restatement of the reference's Triplet A* driver:
//
//   astar/triplet_astar.cpp:991-1622 (astar(): pairwise pass, triples around every variable, unfaithful-edge closure,
//   orientation rules), :811-989 (process_triple), :283-674 (its A* over one cluster), :786-809 (skeleton / cluster update)
//
// The driver consumes the `.pss` only through the best-score structures (spgs[v]->getScore / getParents) and the skeleton,
// all of which search_host.hpp already provides.  What it computes: for every variable i and every pair (vj, vk) of its
// skeleton neighbours, the optimal network over the union of the three clusters (A*, static pattern database); a variable
// that gets both others as parents is a collider and orients two edges, otherwise adjacent pairs become undirected edges;
// edges found outside the skeleton ("unfaithful") are added to it and their neighbourhoods re-examined until nothing new
// appears; three orientation rules then run to a fixed point.  Output: the matrix M with M[i][j] = 1 iff i -> j, both
// directions set for an undirected edge (README.md:40-45).
//
// Behaviours kept because they decide the published outputs (triplet_data/Figure_1, Figure_2):
//   * the neighbour list of i is every set bit of its skeleton row — with a diagonal of ones that includes i itself, and the
//     degenerate triple (i, i, vk) then orients i -> vk whenever the cluster's optimal network has i among vk's parents;
//   * a cluster starts as neighbours + self but is reset to the bare skeleton row when an edge is added at that variable;
//   * triples are de-duplicated by their sorted indices; clusters of more than 26 variables are skipped;
//   * the A* here re-opens a closed node that gets a better g and does not apply the skeleton's leaf filter;
//   * "rule 3" and "rule 4" are the reference's own variants of Meek's rules, as coded at :1334-1425.
#pragma once
#include <set>

#include "search_host.hpp"

namespace urlsearch {

struct TripletResult {
    std::vector<std::vector<int>> directed;   // [i][j] = 1: i -> j; both 1: undirected
    int triplesRun = 0, vStructures = 0, unfaithfulEdges = 0, orientedByRules = 0, ruleIterations = 0;
    long nodesExpanded = 0;
};

class TripletDriver {
public:
    // neighbours[v] = row v of the skeleton as read (diagonal included if the file has one)
    TripletDriver(int variableCount, std::vector<BestScoreCalculator *> &spgs, const std::vector<varset> &neighbours, int pdCount = 2)
        : n(variableCount), spgs(spgs), nb(neighbours), pdCount(pdCount) {
        if (n > 60) throw std::runtime_error("triplet: more than 60 variables");
        if ((int)nb.size() != n) throw std::runtime_error("triplet: the skeleton has a different number of variables than the score file");
        cluster.resize(n);
        for (int v = 0; v < n; v++) cluster[v] = nb[v] | bit(v);               // :1057-1059
        res.directed.assign(n, std::vector<int>(n, 0));
        collider_parents.assign(n, 0);
    }

    TripletResult run() {
        triples_around_every_variable();
        close_under_unfaithful_edges();
        orientation_rules();
        return res;
    }

private:
    static varset bit(int v) { return (varset)1 << v; }
    bool has(varset s, int v) const { return (s >> v) & 1; }
    int &G(int a, int b) { return res.directed[a][b]; }
    bool linked(int a, int b) { return G(a, b) || G(b, a); }

    // the optimal network over one cluster: parents and children of every member (:283-674, :176-235)
    void solve_cluster(varset members, std::vector<varset> &parents, std::vector<varset> &children) {
        StaticPatternDatabase heuristic(n, pdCount, 0, members);
        heuristic.initialize(spgs);
        AstarResult r = run_astar_on_one_scc(n, spgs, heuristic, 0, members, {}, /*reopenClosed=*/true);
        res.nodesExpanded += r.nodesExpanded;
        parents.assign(n, 0);
        children.assign(n, 0);
        if (!r.found) return;
        for (int v = 0; v < n; v++) {
            if (!has(members, v)) continue;
            parents[v] = r.parents[v];
            for (int j = 0; j < n; j++)
                if (has(r.parents[v], j)) children[j] |= bit(v);
        }
    }

    // child <- a, child <- b; the pair (a, b) becomes an undirected edge if the cluster's network joins them and nothing is
    // known about it yet (:876-890 and its two mirror images)
    void mark_collider(int child, int a, int b, const std::vector<varset> &op) {
        res.vStructures++;
        G(a, child) = 1;
        G(b, child) = 1;
        G(child, a) = 0;
        G(child, b) = 0;
        mark_adjacent(a, b, op);
        collider_parents[child] |= bit(a) | bit(b);
    }
    void mark_adjacent(int a, int b, const std::vector<varset> &op) {
        if ((has(op[a], b) || has(op[b], a)) && G(a, b) == 0 && G(b, a) == 0) { G(a, b) = 1; G(b, a) = 1; }
    }

    void process_triple(int i, int vj, int vk) { // :811-989
        const varset big = cluster[i] | cluster[vj] | cluster[vk];
        if (cardinality(big) > 26) return;
        int t[3] = {i, vj, vk};
        std::sort(t, t + 3);
        const uint64_t key = ((uint64_t)t[0] << 40) + ((uint64_t)t[1] << 20) + (uint64_t)t[2];
        if (!done.insert(key).second) return;
        res.triplesRun++;
        std::vector<varset> op, oc;
        solve_cluster(big, op, oc);
        const bool i_is_child = has(op[i], vj) && has(op[i], vk);
        const bool vk_is_child = has(op[vk], i) && has(op[vk], vj);
        const bool vj_is_child = has(op[vj], i) && has(op[vj], vk);
        if (i_is_child) mark_collider(i, vj, vk, op);
        else if (vk_is_child) mark_collider(vk, vj, i, op);       // the reference writes [vj][vk] before [i][vk]; same cells
        else if (vj_is_child) mark_collider(vj, vk, i, op);
        else {
            mark_adjacent(vj, vk, op);
            mark_adjacent(vj, i, op);
            mark_adjacent(vk, i, op);
        }
    }

    void add_skeleton_edge(int a, int b) { // Skeleton::add_edge + the cluster reset of :1196-1198 / :786-809
        nb[a] |= bit(b);
        nb[b] |= bit(a);
        cluster[a] = nb[a];
        cluster[b] = nb[b];
    }

    void triples_around_every_variable() { // :1126-1252
        for (int i = 0; i < n; i++) {
            std::vector<int> around;
            for (int j = 0; j < n; j++)
                if (has(nb[i], j)) around.push_back(j);
            if (around.size() == 1 && i < around[0] && cardinality(nb[around[0]]) == 1) { // an edge with no other neighbour (:1170-1182)
                const int vj = around[0];
                spgs[i]->getScore(bit(vj));
                if (spgs[i]->getParents() == bit(vj)) { G(i, vj) = 1; G(vj, i) = 1; }
            }
            for (size_t a = 0; a < around.size(); a++)
                for (size_t b = 0; b < a; b++) {
                    const int vj = around[a], vk = around[b];
                    process_triple(i, vj, vk);
                    if (!has(nb[vj], vk) && linked(vj, vk)) add_skeleton_edge(vj, vk);
                }
        }
    }

    void close_under_unfaithful_edges() { // :1254-1290
        bool again = true;
        while (again) {
            again = false;
            for (int i = 0; i < n; i++)
                for (int j = 0; j < i; j++) {
                    if (!linked(i, j) || has(nb[i], j)) continue;
                    res.unfaithfulEdges++;
                    again = true;
                    add_skeleton_edge(i, j);
                    for (int k = 0; k < n; k++)
                        if (k != i && k != j && (has(cluster[i], k) || has(cluster[j], k))) process_triple(i, j, k);
                }
        }
    }

    void orientation_rules() { // :1292-1477
        int iter = 0;
        for (; iter < n; iter++) {
            int oriented = 0;
            for (int v = 0; v < n; v++) { // a -> v -> c with a - c undirected: a -> c
                std::vector<int> ins, outs;
                classify(v, &ins, &outs, nullptr);
                if (ins.empty() || outs.empty()) continue;
                for (int a : ins)
                    for (int c : outs)
                        if (G(a, c) && G(c, a)) { G(a, c) = 1; G(c, a) = 0; oriented++; }
            }
            for (int v = 0; v < n; v++) { // v has collider parents; an undirected neighbour tied to two of them points at v
                std::vector<int> undirected;
                classify(v, nullptr, nullptr, &undirected);
                if (cardinality(collider_parents[v]) < 2 || undirected.empty()) continue;
                for (int u : undirected) {
                    int tied = 0;
                    for (int q = 0; q < n; q++)
                        if (has(collider_parents[v], q) && G(q, u) == 1 && G(u, q) == 1) tied++;
                    if (tied >= 2) { G(v, u) = 0; oriented++; }
                }
            }
            for (int v = 0; v < n; v++) { // u - v undirected, u tied to a parent of v: u -> every child of v it is tied to
                std::vector<int> ins, outs, undirected;
                classify(v, &ins, &outs, &undirected);
                if (ins.empty() || outs.empty() || undirected.empty()) continue;
                for (int u : undirected) {
                    bool tied_to_a_parent = false;
                    for (int a : ins)
                        if (G(u, a) == 1 && G(a, u) == 1) tied_to_a_parent = true;
                    if (!tied_to_a_parent) continue;
                    for (int c : outs) {
                        if (G(u, c) == 0 || G(c, u) == 0) continue;
                        oriented++;
                        G(c, u) = 0;
                    }
                }
            }
            res.orientedByRules += oriented;
            if (oriented == 0) break;
        }
        res.ruleIterations = iter;
    }
    // the edges at v as they are when the rule reaches v
    void classify(int v, std::vector<int> *ins, std::vector<int> *outs, std::vector<int> *undirected) {
        for (int j = 0; j < n; j++) {
            const int out = G(v, j), in = G(j, v);
            if (out == 1 && in == 0) { if (outs) outs->push_back(j); }
            else if (in == 1 && out == 0) { if (ins) ins->push_back(j); }
            else if (in == 1 && out == 1) { if (undirected) undirected->push_back(j); }
        }
    }

    int n;
    std::vector<BestScoreCalculator *> &spgs;
    std::vector<varset> nb, cluster, collider_parents;
    int pdCount;
    std::set<uint64_t> done;
    TripletResult res;
};

} // namespace urlsearch
