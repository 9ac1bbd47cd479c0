// One-time symbolic analysis on the host: static pivot plan (maximum-product matching + minimum degree) and
// the fill pattern of the pivot-free LU that every lane then refactors numerically.
//
// The reference re-runs faer's COLAMD + symbolic + partial-pivoting LU at every solve point (op.rs:48-51,
// dc.rs:81-84, ac.rs:67-70, transient.rs:74-77).  Here the pattern is analysed once per circuit x mode and the
// pivot sequence is fixed (KLU-refactor style); MNA's structurally zero diagonals (V sources, inductors:
// stampable.rs:960-983, 676-704) are handled by the row matching.
#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <queue>
#include <set>

#include "plan.hpp"

namespace krb {

// Maximum-product transversal (Duff & Koster MC64 job 5 formulation): minimise sum of
// c_ij = log(max_i |a_ij|) - log |a_ij| over perfect matchings, by successive shortest augmenting paths with
// row/column potentials (reduced costs stay >= 0, matched edges stay tight).
bool weighted_matching(const Pattern &p, const std::vector<double> &mag, std::vector<int> &match_row, double diag_bonus) {
    const int n = p.n;
    const double ZERO_COST = 80.0;   // numerically-zero but structurally present entries: last resort
    std::vector<double> cost(p.nnz);
    for (int j = 0; j < n; j++) {
        double cmax = 0.0;
        for (int q = p.colptr[j]; q < p.colptr[j + 1]; q++) cmax = std::max(cmax, mag[q]);
        for (int q = p.colptr[j]; q < p.colptr[j + 1]; q++)
            cost[q] = (mag[q] > 0.0 && cmax > 0.0) ? std::log(cmax) - std::log(mag[q]) : ZERO_COST;
        if (diag_bonus > 0.0)
            for (int q = p.colptr[j]; q < p.colptr[j + 1]; q++)
                if (p.rowidx[q] == j && cost[q] < ZERO_COST) cost[q] = std::max(0.0, cost[q] - diag_bonus);
    }
    std::vector<double> u(n, 0.0), v(n, 0.0);
    std::vector<int> row_mate(n, -1);   // row -> column
    match_row.assign(n, -1);            // column -> row
    // greedy start: prefer the diagonal when it is (one of) the largest entries of its column, else the largest
    for (int j = 0; j < n; j++) {
        int best = -1;
        double bc = std::numeric_limits<double>::infinity();
        for (int q = p.colptr[j]; q < p.colptr[j + 1]; q++) {
            const int i = p.rowidx[q];
            if (row_mate[i] >= 0) continue;
            if (cost[q] < bc || (cost[q] == bc && i == j)) { bc = cost[q]; best = i; }
        }
        if (best >= 0 && bc == 0.0) { row_mate[best] = j; match_row[j] = best; }
    }
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> d(n), distc(n);
    std::vector<int> pred(n), fin_rows, fin_cols;
    std::vector<char> fin(n, 0), touched(n, 0);
    std::vector<int> touched_list;
    for (int j0 = 0; j0 < n; j0++) {
        if (match_row[j0] >= 0) continue;
        using QE = std::pair<double, int>;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> heap;
        fin_rows.clear(); fin_cols.clear(); touched_list.clear();
        int j = j0;
        double dj = 0.0;
        int free_row = -1;
        double delta = 0.0;
        for (;;) {
            fin_cols.push_back(j);
            distc[j] = dj;
            for (int q = p.colptr[j]; q < p.colptr[j + 1]; q++) {
                const int i = p.rowidx[q];
                if (fin[i]) continue;
                const double nd = dj + (cost[q] - u[i] - v[j]);
                if (!touched[i]) { touched[i] = 1; touched_list.push_back(i); d[i] = INF; }
                if (nd < d[i]) { d[i] = nd; pred[i] = j; heap.push({nd, i}); }
            }
            int i = -1;
            while (!heap.empty()) {
                QE t = heap.top(); heap.pop();
                if (fin[t.second] || t.first > d[t.second]) continue;
                i = t.second;
                break;
            }
            if (i < 0) break;   // no augmenting path: structurally singular
            fin[i] = 1;
            fin_rows.push_back(i);
            if (row_mate[i] < 0) { free_row = i; delta = d[i]; break; }
            j = row_mate[i];
            dj = d[i];
        }
        if (free_row >= 0) {
            for (int jc : fin_cols) v[jc] += delta - distc[jc];
            for (int i : fin_rows) u[i] -= delta - d[i];
            // augment along pred[]
            int i = free_row;
            while (true) {
                const int jc = pred[i];
                const int prev_i = match_row[jc];
                match_row[jc] = i;
                row_mate[i] = jc;
                if (jc == j0) break;
                i = prev_i;
            }
        }
        for (int i : touched_list) { touched[i] = 0; fin[i] = 0; }
        if (free_row < 0) return false;
    }
    return true;
}

// Exact minimum degree on the elimination graph of B + B^T, B = row-permuted A (row j of B = equation matched to
// unknown j).  Ties broken by lowest index.  Ladders/trees come out leaf-first with zero fill.
// exact minimum degree on an explicit (symmetric, sorted, duplicate-free) adjacency; ties broken by lowest index
static void minimum_degree_graph(std::vector<std::vector<int>> &adj, std::vector<int> &order) {
    const int n = (int)adj.size();
    // A path graph (every ladder / transmission line) is eliminated end to end: zero fill, and the permuted matrix is
    // tridiagonal, which the native chain executor relies on.  (Minimum degree gives the same fill but may alternate ends.)
    {
        int leaves = 0, first_leaf = -1;
        bool path = n >= 1;
        for (int i = 0; i < n && path; i++) {
            if (adj[i].size() > 2) path = false;
            else if (adj[i].size() <= 1) { leaves++; if (first_leaf < 0) first_leaf = i; }
        }
        if (path && ((n == 1 && leaves == 1) || (n > 1 && leaves == 2))) {
            order.clear();
            int prev = -1, cur = first_leaf;
            while (cur >= 0 && (int)order.size() < n) {
                order.push_back(cur);
                int nxt = -1;
                for (int w : adj[cur]) if (w != prev) nxt = w;
                prev = cur; cur = nxt;
            }
            if ((int)order.size() == n) return;   // connected path
            order.clear();
        }
    }
    std::set<std::pair<int, int>> pq;
    std::vector<int> deg(n);
    for (int i = 0; i < n; i++) { deg[i] = (int)adj[i].size(); pq.insert({deg[i], i}); }
    std::vector<char> dead(n, 0);
    order.clear();
    order.reserve(n);
    std::vector<int> merged;
    while (!pq.empty()) {
        const int v = pq.begin()->second;
        pq.erase(pq.begin());
        dead[v] = 1;
        order.push_back(v);
        const std::vector<int> nb = adj[v];   // all alive (dead nodes are removed from neighbours' lists below)
        for (int w : nb) {
            // adj[w] = (adj[w] U nb) \ {w, v}
            merged.clear();
            std::set_union(adj[w].begin(), adj[w].end(), nb.begin(), nb.end(), std::back_inserter(merged));
            merged.erase(std::remove_if(merged.begin(), merged.end(), [&](int x) { return x == w || x == v; }), merged.end());
            pq.erase({deg[w], w});
            adj[w].swap(merged);
            deg[w] = (int)adj[w].size();
            pq.insert({deg[w], w});
        }
        std::vector<int>().swap(adj[v]);
    }
}


void minimum_degree(const Pattern &p, const std::vector<int> &match_row, std::vector<int> &order) {
    const int n = p.n;
    std::vector<int> col_of_row(n, -1);
    for (int j = 0; j < n; j++) col_of_row[match_row[j]] = j;
    std::vector<std::vector<int>> adj(n);
    for (int l = 0; l < n; l++)
        for (int q = p.colptr[l]; q < p.colptr[l + 1]; q++) {
            const int j = col_of_row[p.rowidx[q]];
            if (j != l) { adj[j].push_back(l); adj[l].push_back(j); }
        }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    minimum_degree_graph(adj, order);
}

// Constrained variant for the front executor: unknown `follower[k]` (a branch current whose own equation has a structurally
// zero diagonal: V sources stampable.rs:972-980, inductors in DC :688-701) is merged with `leader[k]` (one of its terminal
// nodes) into one vertex of the ordering graph and eliminated right after it, when the node's elimination has filled the
// missing diagonal.  Rows stay in place (identity matching), so the pattern stays structurally symmetric.
void minimum_degree_paired(const Pattern &p, const std::vector<int> &leader_of, std::vector<int> &order) {
    const int n = p.n;
    std::vector<int> rep(n), cid(n, -1);
    for (int u = 0; u < n; u++) rep[u] = leader_of[u] >= 0 ? leader_of[u] : u;
    int nc = 0;
    for (int u = 0; u < n; u++) if (rep[u] == u) cid[u] = nc++;
    std::vector<std::vector<int>> adj(nc);
    for (int col = 0; col < n; col++)
        for (int q = p.colptr[col]; q < p.colptr[col + 1]; q++) {
            const int a = cid[rep[p.rowidx[q]]], b = cid[rep[col]];
            if (a != b) { adj[a].push_back(b); adj[b].push_back(a); }
        }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::vector<int> corder;
    minimum_degree_graph(adj, corder);
    std::vector<std::vector<int>> followers(n);
    for (int u = 0; u < n; u++) if (rep[u] != u) followers[rep[u]].push_back(u);
    std::vector<int> unknown_of_cid(nc);
    for (int u = 0; u < n; u++) if (cid[u] >= 0) unknown_of_cid[cid[u]] = u;
    order.clear();
    for (int c : corder) {
        const int u = unknown_of_cid[c];
        order.push_back(u);
        for (int f : followers[u]) order.push_back(f);
    }
}

// Reverse Cuthill-McKee on B + B^T from a pseudo-peripheral node (candidate order for the banded chain executor).
void rcm_order(const Pattern &p, const std::vector<int> &match_row, std::vector<int> &order) {
    const int n = p.n;
    std::vector<int> col_of_row(n, -1);
    for (int j = 0; j < n; j++) col_of_row[match_row[j]] = j;
    std::vector<std::vector<int>> adj(n);
    for (int l = 0; l < n; l++)
        for (int q = p.colptr[l]; q < p.colptr[l + 1]; q++) {
            const int j = col_of_row[p.rowidx[q]];
            if (j != l) { adj[j].push_back(l); adj[l].push_back(j); }
        }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    for (auto &a : adj) std::sort(a.begin(), a.end(), [&](int x, int z) { return adj[x].size() != adj[z].size() ? adj[x].size() < adj[z].size() : x < z; });
    order.clear();
    std::vector<int> level(n, -1);
    auto bfs = [&](int root, std::vector<int> &out) {
        out.clear();
        out.push_back(root);
        level[root] = 0;
        for (size_t h = 0; h < out.size(); h++)
            for (int w : adj[out[h]]) if (level[w] < 0) { level[w] = level[out[h]] + 1; out.push_back(w); }
    };
    std::vector<char> done(n, 0);
    std::vector<int> comp, comp2;
    for (int s = 0; s < n; s++) {
        if (done[s]) continue;
        int root = s;
        for (int iter = 0; iter < 4; iter++) {          // pseudo-peripheral: restart from a deepest, lowest-degree node
            bfs(root, comp);
            int far = comp.back();
            const int depth = level[far];
            for (int v : comp) if (level[v] == depth && adj[v].size() < adj[far].size()) far = v;
            for (int v : comp) level[v] = -1;
            if (far == root) break;
            bfs(far, comp2);
            const int depth2 = level[comp2.back()];
            for (int v : comp2) level[v] = -1;
            if (depth2 <= depth && iter > 0) break;
            root = far;
        }
        bfs(root, comp);
        for (int v : comp) { done[v] = 1; order.push_back(v); }
    }
    std::reverse(order.begin(), order.end());
}

// lower bandwidth <= 1 and upper bandwidth <= 2 of the row/column-permuted pattern (the class is closed under LU fill)
bool band_ok(const Pattern &p, const std::vector<int> &match_row, const std::vector<int> &order) {
    const int n = p.n;
    std::vector<int> pos_of_col(n), pos_of_row(n);
    for (int k = 0; k < n; k++) { pos_of_col[order[k]] = k; pos_of_row[match_row[order[k]]] = k; }
    for (int col = 0; col < n; col++)
        for (int q = p.colptr[col]; q < p.colptr[col + 1]; q++) {
            const int d = pos_of_col[col] - pos_of_row[p.rowidx[q]];   // > 0: upper part
            if (d < -1 || d > 2) return false;
        }
    return true;
}

void symbolic_lu(const Pattern &p, const std::vector<int> &match_row, const std::vector<int> &order, Symbolic &s) {
    const int n = p.n;
    s.n = n;
    s.col_order = order;
    s.row_order.resize(n);
    std::vector<int> pos_of_col(n);
    for (int k = 0; k < n; k++) { s.row_order[k] = match_row[order[k]]; pos_of_col[order[k]] = k; }
    std::vector<int> pos_of_row(n);
    for (int k = 0; k < n; k++) pos_of_row[s.row_order[k]] = k;
    // rows of the permuted matrix C: (l, entry id)
    std::vector<std::vector<std::pair<int, int>>> crow(n);
    for (int col = 0; col < n; col++)
        for (int q = p.colptr[col]; q < p.colptr[col + 1]; q++) crow[pos_of_row[p.rowidx[q]]].push_back({pos_of_col[col], q});
    s.Lpat.assign(n, {}); s.Upat.assign(n, {}); s.Lent.assign(n, {}); s.Uent.assign(n, {}); s.Dent.assign(n, -1);
    std::vector<int> mark(n, -1), ent(n, -1);
    s.nnz_lu = 0;
    s.flops = 0;
    for (int k = 0; k < n; k++) {
        std::priority_queue<int, std::vector<int>, std::greater<int>> heap;
        std::vector<int> lower, upper;
        bool has_diag = false;
        for (auto &le : crow[k]) {
            const int l = le.first;
            mark[l] = k; ent[l] = le.second;
            if (l < k) heap.push(l);
            else if (l > k) upper.push_back(l);
            else has_diag = true;
        }
        (void)has_diag;
        while (!heap.empty()) {
            const int j = heap.top(); heap.pop();
            lower.push_back(j);
            for (int l : s.Upat[j]) {
                s.flops += 2;
                if (mark[l] == k) continue;
                mark[l] = k; ent[l] = -1;
                if (l < k) heap.push(l);
                else if (l > k) upper.push_back(l);
            }
            s.flops += 2;   // y update
        }
        std::sort(upper.begin(), upper.end());
        s.Lpat[k] = lower;
        s.Upat[k] = upper;
        for (int j : lower) s.Lent[k].push_back(ent[j]);
        for (int l : upper) s.Uent[k].push_back(ent[l]);
        s.Dent[k] = (mark[k] == k) ? ent[k] : -1;
        if (mark[k] != k) { mark[k] = k; ent[k] = -1; }   // structurally zero diagonal that only fill can populate
        s.nnz_lu += (int64_t)lower.size() + (int64_t)upper.size() + 1;
        s.flops += 2 * (int64_t)upper.size() + 4;          // row scaling + back substitution
    }
}

}  // namespace krb
