// Host symbolic analysis for the front executor: elimination tree, relaxed supernodes and the assembly (front) tree of
// the symmetrised, statically pivoted matrix.  Runs once per circuit x mode (the work the reference redoes inside faer's
// sp_lu() at every solve point: op.rs:50, dc.rs:83, ac.rs:70, transient.rs:76).
//
// Input: the stamped CSC pattern, the row matching (static pivots) and a fill-reducing elimination order of the unknowns.
// Output (FrontPlan): fronts = dense (p + q) x (p + q) blocks, p pivots eliminated in the front, q update indices passed on
// to the parent as a contribution block.  Fronts of one level of the tree are independent and are factorised by one launch.
#include <algorithm>
#include <cassert>
#include <cstdlib>
#include <stdexcept>

#include "plan.hpp"

namespace krb {

namespace {

// struct_[k] = row indices > k of column k of L (symmetrised pattern), sorted; parent[k] = min(struct_[k]) or -1
void symbolic_columns(int n, const std::vector<std::vector<int>> &up, std::vector<std::vector<int>> &st, std::vector<int> &parent) {
    st.assign(n, {});
    parent.assign(n, -1);
    std::vector<std::vector<int>> kids(n);
    std::vector<int> mark(n, -1);
    for (int k = 0; k < n; k++) {
        std::vector<int> &s = st[k];
        mark[k] = k;
        for (int v : up[k]) if (mark[v] != k) { mark[v] = k; s.push_back(v); }
        for (int c : kids[k])
            for (int v : st[c]) if (mark[v] != k) { mark[v] = k; s.push_back(v); }
        std::sort(s.begin(), s.end());
        if (!s.empty()) { parent[k] = s[0]; kids[s[0]].push_back(k); }
    }
}

}  // namespace

void build_front_plan(const Pattern &P, const std::vector<int> &match_row, const std::vector<int> &order, FrontPlan &fp) {
    const int n = P.n;
    fp = FrontPlan();
    fp.n = n;
    if (n == 0) { fp.valid = true; return; }
    // ---- pass 1: structure in the given order, then postorder the elimination tree ----
    std::vector<int> pos_of_col(n), pos_of_row(n);
    for (int k = 0; k < n; k++) { pos_of_col[order[k]] = k; pos_of_row[match_row[order[k]]] = k; }
    auto upper_adj = [&](std::vector<std::vector<int>> &up) {
        up.assign(n, {});
        for (int col = 0; col < n; col++)
            for (int e = P.colptr[col]; e < P.colptr[col + 1]; e++) {
                const int r = pos_of_row[P.rowidx[e]], c = pos_of_col[col];
                if (r != c) up[std::min(r, c)].push_back(std::max(r, c));
            }
        for (auto &u : up) { std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end()); }
    };
    std::vector<std::vector<int>> up, st;
    std::vector<int> parent;
    upper_adj(up);
    symbolic_columns(n, up, st, parent);
    // postorder (children in ascending order, iterative DFS from the roots)
    std::vector<int> post(n, -1);
    {
        std::vector<std::vector<int>> kids(n);
        std::vector<int> roots;
        for (int k = 0; k < n; k++) (parent[k] >= 0 ? kids[parent[k]] : roots).push_back(k);
        int next = 0;
        std::vector<std::pair<int, size_t>> stack;
        for (int r : roots) {
            stack.push_back({r, 0});
            while (!stack.empty()) {
                auto &top = stack.back();
                if (top.second < kids[top.first].size()) { const int c = kids[top.first][top.second++]; stack.push_back({c, 0}); }
                else { post[top.first] = next++; stack.pop_back(); }
            }
        }
    }
    // relabel: structure sets and parents in postorder numbering (same filled graph: a postorder is an equivalent ordering)
    std::vector<int> ord1(n);
    for (int k = 0; k < n; k++) ord1[post[k]] = order[k];
    {
        std::vector<std::vector<int>> st2(n);
        std::vector<int> par2(n, -1);
        for (int k = 0; k < n; k++) {
            std::vector<int> &d = st2[post[k]];
            d.reserve(st[k].size());
            for (int v : st[k]) d.push_back(post[v]);
            std::sort(d.begin(), d.end());
            par2[post[k]] = parent[k] >= 0 ? post[parent[k]] : -1;
            std::vector<int>().swap(st[k]);
        }
        st.swap(st2);
        parent.swap(par2);
    }
    // ---- fundamental supernodes (exact): k joins k-1 when column k-1's structure is {k} + column k's ----
    std::vector<int> sn_of(n);
    std::vector<int> sn_first, sn_last;
    for (int k = 0; k < n; k++) {
        const bool join = k > 0 && parent[k - 1] == k && st[k - 1].size() == st[k].size() + 1;
        if (!join) { sn_first.push_back(k); sn_last.push_back(k); }
        else sn_last.back() = k;
        sn_of[k] = (int)sn_first.size() - 1;
    }
    const int ns = (int)sn_first.size();
    std::vector<int> sn_parent(ns, -1);
    for (int s = 0; s < ns; s++) { const int pk = parent[sn_last[s]]; sn_parent[s] = pk >= 0 ? sn_of[pk] : -1; }
    // ---- relaxed amalgamation (CHOLMOD-style thresholds): merge a child into its parent when the explicit zeros stay small ----
    int relax_p = 16;
    if (const char *e = getenv("KRB_FRONT_RELAX")) relax_p = std::max(1, atoi(e));
    struct SN { std::vector<int> piv; int q; int64_t zeros; bool alive; std::vector<int> kids; };
    std::vector<SN> sn(ns);
    for (int s = 0; s < ns; s++) {
        for (int k = sn_first[s]; k <= sn_last[s]; k++) sn[s].piv.push_back(k);
        sn[s].q = (int)st[sn_last[s]].size();
        sn[s].zeros = 0;
        sn[s].alive = true;
    }
    for (int s = 0; s < ns; s++) if (sn_parent[s] >= 0) sn[sn_parent[s]].kids.push_back(s);
    for (int s = 0; s < ns; s++) {   // ascending = children before parents
        SN &me = sn[s];
        std::vector<int> kids = me.kids;
        std::sort(kids.begin(), kids.end(), [&](int a, int b) { return sn[a].piv.size() + sn[a].q != sn[b].piv.size() + sn[b].q ? sn[a].piv.size() + sn[a].q > sn[b].piv.size() + sn[b].q : a < b; });
        std::vector<int> keep;
        std::vector<std::vector<int>> merged_piv;
        for (int c : kids) {
            SN &ch = sn[c];
            const int64_t pc = (int64_t)ch.piv.size();
            int64_t pm = (int64_t)me.piv.size();
            for (auto &mp : merged_piv) pm += (int64_t)mp.size();
            const int64_t p_new = pc + pm;
            auto stored = [](int64_t p, int64_t q) { return p * (p + 2 * q); };   // m x p block + p x q block
            const int64_t z_new = stored(p_new, me.q) - stored(pm, me.q) - stored(pc, ch.q);   // explicit zeros this merge adds
            const int64_t z_tot = z_new + ch.zeros + me.zeros;
            const int64_t size_tot = stored(p_new, me.q);
            const double frac = size_tot > 0 ? (double)z_tot / (double)size_tot : 0.0;
            bool merge = false;
            // every front costs a CTA (or warp) with a fixed chain of dependent global accesses, so small fronts are merged
            // generously; larger ones only when the explicit zeros stay a small fraction of the stored block
            if (z_new == 0) merge = true;
            else if (p_new <= relax_p / 4) merge = true;
            else if (p_new <= relax_p) merge = frac < 0.8;
            else if (p_new <= 3 * relax_p) merge = frac < 0.1;
            else merge = frac < 0.05;
            if (merge) {
                merged_piv.push_back(ch.piv);
                me.zeros += z_new + ch.zeros;
                for (int g : ch.kids) keep.push_back(g);   // grandchildren move up
                ch.alive = false;
                std::vector<int>().swap(ch.piv);
            } else keep.push_back(c);
        }
        if (!merged_piv.empty()) {
            std::vector<int> np;
            for (auto &mp : merged_piv) np.insert(np.end(), mp.begin(), mp.end());
            np.insert(np.end(), me.piv.begin(), me.piv.end());
            me.piv.swap(np);
        }
        std::sort(keep.begin(), keep.end());
        me.kids.swap(keep);
    }
    // ---- final numbering: postorder of the amalgamated tree, each front's pivots contiguous ----
    std::vector<int> newpos(n, -1), front_of_sn(ns, -1);
    std::vector<int> f_sn;   // front -> supernode
    {
        std::vector<int> roots;
        std::vector<char> is_child(ns, 0);
        for (int s = 0; s < ns; s++) if (sn[s].alive) for (int c : sn[s].kids) is_child[c] = 1;
        for (int s = 0; s < ns; s++) if (sn[s].alive && !is_child[s]) roots.push_back(s);
        int next = 0;
        std::vector<std::pair<int, size_t>> stack;
        for (int r : roots) {
            stack.push_back({r, 0});
            while (!stack.empty()) {
                auto &top = stack.back();
                if (top.second < sn[top.first].kids.size()) { const int c = sn[top.first].kids[top.second++]; stack.push_back({c, 0}); }
                else {
                    const int s = top.first;
                    front_of_sn[s] = (int)f_sn.size();
                    f_sn.push_back(s);
                    for (int k : sn[s].piv) newpos[k] = next++;
                    stack.pop_back();
                }
            }
        }
        if (next != n) throw std::runtime_error("front plan: internal numbering error");
    }
    const int nf = (int)f_sn.size();
    fp.n_fronts = nf;
    fp.row_order.resize(n);
    fp.col_order.resize(n);
    for (int k = 0; k < n; k++) { fp.col_order[newpos[k]] = ord1[k]; fp.row_order[newpos[k]] = match_row[ord1[k]]; }
    fp.piv0.assign(nf + 1, 0);
    for (int f = 0; f < nf; f++) fp.piv0[f + 1] = fp.piv0[f] + (int)sn[f_sn[f]].piv.size();
    std::vector<int> front_of_pos(n);
    for (int f = 0; f < nf; f++) for (int k = fp.piv0[f]; k < fp.piv0[f + 1]; k++) front_of_pos[k] = f;
    // update sets: union of the pivots' column structures, minus the pivots
    fp.idx_ptr.assign(nf + 1, 0);
    fp.parent.assign(nf, -1);
    {
        std::vector<int> mark(n, -1);
        for (int f = 0; f < nf; f++) {
            std::vector<int> u;
            const int lo = fp.piv0[f], hi = fp.piv0[f + 1];
            for (int k : sn[f_sn[f]].piv)
                for (int v : st[k]) {
                    const int w = newpos[v];
                    if (w >= lo && w < hi) continue;
                    if (w < lo) throw std::runtime_error("front plan: update index precedes its front");
                    if (mark[w] != f) { mark[w] = f; u.push_back(w); }
                }
            std::sort(u.begin(), u.end());
            fp.idx.insert(fp.idx.end(), u.begin(), u.end());
            fp.idx_ptr[f + 1] = (int)fp.idx.size();
            if (!u.empty()) fp.parent[f] = front_of_pos[u[0]];
        }
    }
    std::vector<std::vector<int>>().swap(st);
    // children, relative indices
    fp.child_ptr.assign(nf + 1, 0);
    for (int f = 0; f < nf; f++) if (fp.parent[f] >= 0) fp.child_ptr[fp.parent[f] + 1]++;
    for (int f = 0; f < nf; f++) fp.child_ptr[f + 1] += fp.child_ptr[f];
    fp.child.assign(fp.child_ptr[nf], -1);
    {
        std::vector<int> fill(fp.child_ptr.begin(), fp.child_ptr.end() - 1);
        for (int f = 0; f < nf; f++) if (fp.parent[f] >= 0) fp.child[fill[fp.parent[f]]++] = f;
    }
    auto local_index = [&](int f, int pos) -> int {   // local row/col of position `pos` inside front f, -1 if absent
        const int lo = fp.piv0[f], hi = fp.piv0[f + 1];
        if (pos >= lo && pos < hi) return pos - lo;
        auto b = fp.idx.begin() + fp.idx_ptr[f], e = fp.idx.begin() + fp.idx_ptr[f + 1];
        auto it = std::lower_bound(b, e, pos);
        if (it == e || *it != pos) return -1;
        return (hi - lo) + (int)(it - b);
    };
    fp.rel_ptr = fp.idx_ptr;
    fp.rel.assign(fp.idx.size(), -1);
    for (int f = 0; f < nf; f++) {
        const int pf = fp.parent[f];
        for (int t = fp.idx_ptr[f]; t < fp.idx_ptr[f + 1]; t++) {
            const int li = local_index(pf, fp.idx[t]);
            if (li < 0) throw std::runtime_error("front plan: child update index missing from its parent front");
            fp.rel[t] = li;
        }
    }
    // assembly lists of the stamped entries
    {
        std::vector<int> npos_of_col(n), npos_of_row(n);
        for (int k = 0; k < n; k++) { npos_of_col[fp.col_order[k]] = k; npos_of_row[fp.row_order[k]] = k; }
        std::vector<std::vector<std::pair<int, int>>> per_front(nf);
        for (int col = 0; col < n; col++)
            for (int e = P.colptr[col]; e < P.colptr[col + 1]; e++) {
                const int r = npos_of_row[P.rowidx[e]], c = npos_of_col[col];
                const int f = front_of_pos[std::min(r, c)];
                const int m = (fp.piv0[f + 1] - fp.piv0[f]) + (fp.idx_ptr[f + 1] - fp.idx_ptr[f]);
                const int lr = local_index(f, r), lc = local_index(f, c);
                if (lr < 0 || lc < 0) throw std::runtime_error("front plan: stamped entry outside its front");
                per_front[f].push_back({e, lr + m * lc});
            }
        fp.asm_ptr.assign(nf + 1, 0);
        for (int f = 0; f < nf; f++) {
            // sorted by position (column-major): the left-looking kernel assembles a front one column block at a time
            std::sort(per_front[f].begin(), per_front[f].end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second < b.second; });
            for (auto &pr : per_front[f]) { fp.asm_ent.push_back(pr.first); fp.asm_pos.push_back(pr.second); }
            fp.asm_ptr[f + 1] = (int)fp.asm_ent.size();
        }
    }
    // levels (leaves first), offsets, counts
    std::vector<int> level(nf, 0);
    int nlev = 0;
    for (int f = 0; f < nf; f++) {
        if (fp.parent[f] >= 0) level[fp.parent[f]] = std::max(level[fp.parent[f]], level[f] + 1);   // children precede parents
        nlev = std::max(nlev, level[f] + 1);
    }
    fp.n_levels = nlev;
    fp.level_ptr.assign(nlev + 1, 0);
    for (int f = 0; f < nf; f++) fp.level_ptr[level[f] + 1]++;
    for (int l = 0; l < nlev; l++) fp.level_ptr[l + 1] += fp.level_ptr[l];
    fp.level_fronts.assign(nf, -1);
    {
        std::vector<int> fill(fp.level_ptr.begin(), fp.level_ptr.end() - 1);
        for (int f = 0; f < nf; f++) fp.level_fronts[fill[level[f]]++] = f;
        auto msize = [&](int f) { return (fp.piv0[f + 1] - fp.piv0[f]) + (fp.idx_ptr[f + 1] - fp.idx_ptr[f]); };
        for (int l = 0; l < nlev; l++)
            std::stable_sort(fp.level_fronts.begin() + fp.level_ptr[l], fp.level_fronts.begin() + fp.level_ptr[l + 1], [&](int a, int b) { return msize(a) > msize(b); });
    }
    fp.lu_off.assign(nf, 0); fp.cb_off.assign(nf, 0); fp.cbv_off.assign(nf, 0);
    // storage in level order, so that the fronts one launch touches are neighbours in memory
    for (int t = 0; t < nf; t++) {
        const int f = fp.level_fronts[t];
        const int64_t p = fp.piv0[f + 1] - fp.piv0[f], q = fp.idx_ptr[f + 1] - fp.idx_ptr[f], m = p + q;
        fp.lu_off[f] = fp.lu_size; fp.lu_size += ((m * p + p * q + 1) / 2) * 2;   // 16-byte aligned blocks (f64)
        fp.cb_off[f] = fp.cb_size; fp.cb_size += ((q * q + 1) / 2) * 2;
        fp.cbv_off[f] = fp.cbv_size; fp.cbv_size += q;
        fp.nnz_lu += m * p + p * q;
        fp.max_m = std::max(fp.max_m, (int)m);
        fp.max_p = std::max(fp.max_p, (int)p);
        for (int64_t k = 0; k < p; k++) fp.flops += (m - k - 1) + 2 * (m - k - 1) * (m - k - 1);
        fp.flops += 2 * (m * p + p * q);
    }
    fp.valid = true;
}

}  // namespace krb
