// Plan compiler: symbolic stamps + static pivot plan  ->  straight-line lane program (vm_ops.h).
//
// The numeric method is a row-wise (up-looking) Crout/LDU factorisation without pivoting,  A = L~ * U^  with U^ unit
// upper triangular, fused with the forward substitution; the back substitution follows in the same program:
//     row k:   w   = A[k,:]                                  (MNA stamping, contributions summed in stamp order)
//              w  -= w[j] * U^[j,:]      for j in Lpat[k]    (y_k -= w[j] * y^_j alongside)
//              U^[k,l] = w[l] / w[k],  y^_k = y_k / w[k]
//     back:    x_k = y^_k - sum_l U^[k,l] x_l
// Every value is a "cell"; the allocator below maps cells to per-lane on-chip scratch slots with Belady
// (farthest-next-use) eviction, rematerialises constants/inputs instead of spilling them, and spills the rest to
// HBM value planes in production order (so the fused back substitution re-reads them as a reverse stream).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <complex>
#include <cstring>
#include <map>
#include <tuple>

#include "plan.hpp"

namespace krb {

namespace {

constexpr uint8_t VOP_PERSIST = 0xFF;   // marker: cell `a` is final and must survive the section

struct VOp {
    uint8_t opc;
    uint8_t flags;
    int d, a, b;
    uint32_t imm;
};
enum Remat : uint8_t { R_NONE = 0, R_LDB = 1, R_LDI = 2 };
struct Cell {
    Remat remat = R_NONE;
    uint32_t rimm = 0;
    int home = -1;        // HBM plane
    bool persist = false; // produced in the factor-once section, consumed in the per-step section
    int pin_slot = -1;
};

struct Builder {
    Plan &plan;
    std::vector<Cell> cells;
    std::vector<int> input_of_elem;
    std::map<std::tuple<int, int, int, uint64_t>, int> pool_index;   // (kind, neg, elem, bits(c)) -> pool idx
    std::map<int, int> pool_cells;                                   // pool idx -> shared remat cell
    std::map<int, int> diode_pool_base;

    explicit Builder(Plan &p) : plan(p) {}

    int new_cell() { cells.emplace_back(); return (int)cells.size() - 1; }
    int pool_idx(PoolKind k, bool neg, int elem, double c = 0.0) {
        uint64_t bits;
        std::memcpy(&bits, &c, 8);
        auto key = std::make_tuple((int)k, (int)neg, elem, bits);
        auto it = pool_index.find(key);
        if (it != pool_index.end()) return it->second;
        plan.pool.push_back({k, neg, elem, c});
        int idx = (int)plan.pool.size() - 1;
        pool_index[key] = idx;
        return idx;
    }
    int shared_pool_cell(int idx) {
        auto it = pool_cells.find(idx);
        if (it != pool_cells.end()) return it->second;
        int c = new_cell();
        cells[c].remat = R_LDB; cells[c].rimm = (uint32_t)idx;
        pool_cells[idx] = c;
        return c;
    }
    int fresh_pool_cell(int idx) {
        int c = new_cell();
        cells[c].remat = R_LDB; cells[c].rimm = (uint32_t)idx;
        return c;
    }
    int fresh_input_cell(int plane) {
        int c = new_cell();
        cells[c].remat = R_LDI; cells[c].rimm = (uint32_t)plane;
        return c;
    }
    int diode_base(int e) {
        auto it = diode_pool_base.find(e);
        if (it != diode_pool_base.end()) return it->second;
        // four consecutive entries: {is/(n*Vt), n*Vt, is, vcrit}
        plan.pool.push_back({P_DIO_ISNVT, false, e, 0.0});
        int base = (int)plan.pool.size() - 1;
        plan.pool.push_back({P_DIO_NVT, false, e, 0.0});
        plan.pool.push_back({P_DIO_IS, false, e, 0.0});
        plan.pool.push_back({P_DIO_VCRIT, false, e, 0.0});
        diode_pool_base[e] = base;
        return base;
    }
    std::map<int, int> mos_pool_base;
    int mos_base(int e) {
        auto it = mos_pool_base.find(e);
        if (it != mos_pool_base.end()) return it->second;
        // three consecutive entries: {Vth, beta, lambda}
        plan.pool.push_back({P_MOS_VTH, false, e, 0.0});
        int base = (int)plan.pool.size() - 1;
        plan.pool.push_back({P_MOS_BETA, false, e, 0.0});
        plan.pool.push_back({P_MOS_LAMBDA, false, e, 0.0});
        mos_pool_base[e] = base;
        return base;
    }
};

// A source reference that can be materialised either as a shared read-only cell or a fresh mutable one
struct Section {
    Builder &B;
    std::vector<VOp> vops;
    int ldx_which_guess;   // LDX flag for "the linearisation point": 1 = NR guess (DC/TRAN), 0 = prev row (AC: OP solution)
    std::map<std::tuple<int, int, int>, int> derived;   // (tag, elem, neg) -> shared cell, valid inside this section
    std::map<int, int> input_cells;                    // plane -> shared remat cell
    // linear transient, everything on chip: previous-row values and step-invariant companion coefficients are carried in
    // pinned scratch cells instead of being re-read from HBM / re-derived at every step
    std::map<int, int> *carry = nullptr;               // result-row column -> persistent cell holding its previous-row value
    std::map<int, int> *coef_voh = nullptr;            // element -> persistent cell holding value / h (varied elements)

    explicit Section(Builder &b, int which_guess) : B(b), ldx_which_guess(which_guess) {}

    int op(uint8_t opc, int d, int a = -1, int b = -1, uint32_t imm = 0, uint8_t flags = 0) {
        vops.push_back({opc, flags, d, a, b, imm});
        return d;
    }
    int un(uint8_t opc, int a, uint32_t imm = 0) { return op(opc, B.new_cell(), a, -1, imm); }
    int bin(uint8_t opc, int a, int b) { return op(opc, B.new_cell(), a, b); }
    int shared_input(int plane) {
        auto it = input_cells.find(plane);
        if (it != input_cells.end()) return it->second;
        int c = B.fresh_input_cell(plane);
        input_cells[plane] = c;
        return c;
    }
    int ldx(int which, int unknown, bool half_default) {
        if (carry && which == 0) {
            auto it = carry->find(unknown);
            if (it != carry->end()) return it->second;
        }
        return op(KOP_LDX, B.new_cell(), -1, -1, (uint32_t)unknown, (uint8_t)((which & 1) | (half_default ? 2 : 0)));
    }
    enum Tag { T_G = 1, T_VAL, T_VOH, T_TWOPIF, T_JWV, T_VD, T_DIOG, T_DIOI, T_CAPH, T_INDH, T_CAPI, T_INDV, T_MOSVGS, T_MOSVDS, T_MOSGM, T_MOSGDS, T_MOSGDSGM, T_MOSIEQ };
    int cached(Tag t, int e, bool neg, int (Section::*make)(int)) {
        auto key = std::make_tuple((int)t, e, (int)neg);
        auto it = derived.find(key);
        if (it != derived.end()) return it->second;
        int c;
        if (neg) c = un(KOP_NEG, cached(t, e, false, make));
        else c = (this->*make)(e);
        derived[key] = c;
        return c;
    }
    bool varied(int e) const { return B.input_of_elem[e] >= 0; }
    int elem_value(int e) { return varied(e) ? shared_input(B.input_of_elem[e]) : B.shared_pool_cell(B.pool_idx(P_VAL, false, e)); }
    int make_g(int e) { return un(KOP_INV, elem_value(e)); }                                   // 1.0 / value
    int h_cell() { return B.shared_pool_cell(B.pool_idx(P_CONST, false, -1, B.plan.h_comp)); }
    int make_voh(int e) {
        if (coef_voh && varied(e)) { auto it = coef_voh->find(e); if (it != coef_voh->end()) return it->second; }
        return make_voh_raw(e);
    }
    int make_voh_raw(int e) { return varied(e) ? bin(KOP_DIVR, elem_value(e), h_cell()) : B.shared_pool_cell(B.pool_idx(P_VOH, false, e)); }
    int make_twopif(int) {
        const double two_pi = 2.0 * 3.14159265358979323846264338327950288;                     // 2.0 * PI
        return bin(KOP_MUL, B.shared_pool_cell(B.pool_idx(P_CONST, false, -1, two_pi)), shared_input(B.plan.input_freq));
    }
    int make_jwv(int e) {                                                                      // (0, ((2.0*PI)*f)*value)
        int w = cached(T_TWOPIF, -1, false, &Section::make_twopif);
        return un(KOP_MULJ, bin(KOP_MUL, w, elem_value(e)));
    }
    int make_vd(int e) {                                                                       // diode.rs:26-40
        const krb_element &el = B.plan.circuit.el[e];
        int vp = el.ip >= 0 ? ldx(ldx_which_guess, el.ip, true) : B.shared_pool_cell(B.pool_idx(P_CONST, false, -1, 0.5));
        if (el.im < 0) return vp;                                                              // vp - 0.0 == vp
        int vm = ldx(ldx_which_guess, el.im, false);
        return bin(KOP_SUB, vp, vm);
    }
    int make_diog(int e) { return un(KOP_DIOG, cached(T_VD, e, false, &Section::make_vd), (uint32_t)B.diode_base(e)); }
    int make_dioi(int e) { return un(KOP_DIOI, cached(T_VD, e, false, &Section::make_vd), (uint32_t)B.diode_base(e)); }
    // level-1 NMOS (extension, KRB_ELEM_FLAG_LEVEL1): node voltages of the linearisation point default to 0.0 (stampable.rs:825-833)
    int mos_node(int u) { return u >= 0 ? ldx(ldx_which_guess, u, false) : B.shared_pool_cell(B.pool_idx(P_CONST, false, -1, 0.0)); }
    int make_mos_vgs(int e) { const krb_element &el = B.plan.circuit.el[e]; return bin(KOP_SUB, mos_node(el.im), mos_node(el.ic)); }   // v_g - v_s
    int make_mos_vds(int e) { const krb_element &el = B.plan.circuit.el[e]; return bin(KOP_SUB, mos_node(el.ip), mos_node(el.ic)); }   // v_d - v_s
    int mos_op(uint8_t opc, int e) {
        return op(opc, B.new_cell(), cached(T_MOSVGS, e, false, &Section::make_mos_vgs), cached(T_MOSVDS, e, false, &Section::make_mos_vds), (uint32_t)B.mos_base(e));
    }
    int make_mos_gm(int e) { return mos_op(KOP_MOSGM, e); }
    int make_mos_gds(int e) { return mos_op(KOP_MOSGDS, e); }
    int make_mos_gdsgm(int e) { return bin(KOP_ADD, cached(T_MOSGDS, e, false, &Section::make_mos_gds), cached(T_MOSGM, e, false, &Section::make_mos_gm)); }
    int make_mos_ieq(int e) {                                                                  // i_d - g_ds * v_ds - g_m * v_gs  (stampable.rs:889)
        const int vgs = cached(T_MOSVGS, e, false, &Section::make_mos_vgs), vds = cached(T_MOSVDS, e, false, &Section::make_mos_vds);
        const int t1 = bin(KOP_MUL, cached(T_MOSGDS, e, false, &Section::make_mos_gds), vds);
        const int t2 = bin(KOP_MUL, cached(T_MOSGM, e, false, &Section::make_mos_gm), vgs);
        return bin(KOP_SUB, bin(KOP_SUB, mos_op(KOP_MOSID, e), t1), t2);
    }
    int make_caph(int e) {                                                                     // (value/h) * (v+_prev - v-_prev)
        const krb_element &el = B.plan.circuit.el[e];
        int coef = varied(e) ? cached(T_VOH, e, false, &Section::make_voh) : make_voh(e);
        int v;
        if (el.ip >= 0 && el.im >= 0) v = bin(KOP_SUB, ldx(0, el.ip, false), ldx(0, el.im, false));
        else if (el.ip >= 0) v = ldx(0, el.ip, false);
        else v = un(KOP_NEG, ldx(0, el.im, false));
        return bin(KOP_MUL, coef, v);
    }
    int make_capi(int e) {                                                                     // trapezoidal: i_C of the previous row
        auto &se = B.plan.state_elems;
        const int j = (int)(std::find(se.begin(), se.end(), e) - se.begin());
        return ldx(0, B.plan.circuit.n + j, false);
    }
    int make_indv(int e) {                                                                     // v+_prev - v-_prev across the inductor
        const krb_element &el = B.plan.circuit.el[e];
        if (el.ip >= 0 && el.im >= 0) return bin(KOP_SUB, ldx(0, el.ip, false), ldx(0, el.im, false));
        if (el.ip >= 0) return ldx(0, el.ip, false);
        if (el.im >= 0) return un(KOP_NEG, ldx(0, el.im, false));
        return op(KOP_ZERO, B.new_cell());
    }
    int make_indh(int e) {                                                                     // (value/h) * i_prev
        const krb_element &el = B.plan.circuit.el[e];
        int coef = varied(e) ? cached(T_VOH, e, false, &Section::make_voh) : make_voh(e);
        return bin(KOP_MUL, coef, ldx(0, el.ic, false));
    }

    // Returns a cell holding the contribution.  `fresh` is set when the cell is private (safe to mutate in place).
    int src_cell(const Src &s, bool &fresh) {
        fresh = false;
        const int e = s.elem;
        switch (s.kind) {
        case S_ONE: fresh = true; return B.fresh_pool_cell(B.pool_idx(P_CONST, s.neg, -1, 1.0));
        case S_G:
            if (!varied(e)) { fresh = true; return B.fresh_pool_cell(B.pool_idx(P_G, s.neg, e)); }
            return cached(T_G, e, s.neg, &Section::make_g);
        case S_VAL:
            if (!varied(e)) { fresh = true; return B.fresh_pool_cell(B.pool_idx(P_VAL, s.neg, e)); }
            if (!s.neg) { fresh = true; return B.fresh_input_cell(B.input_of_elem[e]); }
            return cached(T_VAL, e, true, &Section::elem_value_wrap);
        case S_VOH:
            if (!varied(e)) { fresh = true; return B.fresh_pool_cell(B.pool_idx(P_VOH, s.neg, e)); }
            return cached(T_VOH, e, s.neg, &Section::make_voh);
        case S_JWV: return cached(T_JWV, e, s.neg, &Section::make_jwv);
        case S_ACAMP: fresh = true; return B.fresh_pool_cell(B.pool_idx(P_ACAMP, s.neg, e));
        case S_WAVE: {
            auto &we = B.plan.wave_elems;
            int col = (int)(std::find(we.begin(), we.end(), e) - we.begin());
            if (col == (int)we.size()) we.push_back(e);
            int c = op(KOP_LDBS, B.new_cell(), -1, -1, (uint32_t)col);
            fresh = true;
            return s.neg ? un(KOP_NEG, c) : c;
        }
        case S_DIODE_G: return cached(T_DIOG, e, s.neg, &Section::make_diog);
        case S_DIODE_IEQ: return cached(T_DIOI, e, s.neg, &Section::make_dioi);
        case S_MOS_GDS: return cached(T_MOSGDS, e, s.neg, &Section::make_mos_gds);
        case S_MOS_GM: return cached(T_MOSGM, e, s.neg, &Section::make_mos_gm);
        case S_MOS_GDSGM: return cached(T_MOSGDSGM, e, s.neg, &Section::make_mos_gdsgm);
        case S_MOS_IEQ: return cached(T_MOSIEQ, e, s.neg, &Section::make_mos_ieq);
        case S_CAP_HIST: return cached(T_CAPH, e, s.neg, &Section::make_caph);
        case S_IND_HIST: return cached(T_INDH, e, s.neg, &Section::make_indh);
        case S_CAP_IPREV: return cached(T_CAPI, e, s.neg, &Section::make_capi);
        case S_IND_VPREV: return cached(T_INDV, e, s.neg, &Section::make_indv);
        }
        return -1;
    }
    int elem_value_wrap(int e) { return elem_value(e); }

    // sum of contributions in stamp order (sum_triplets, solver.rs:59-70); result is a private mutable cell
    int sum_cell(const std::vector<Src> &srcs) {
        if (srcs.empty()) return op(KOP_ZERO, B.new_cell());
        bool fresh;
        int acc = src_cell(srcs[0], fresh);
        if (srcs.size() == 1) return fresh ? acc : un(KOP_MOV, acc);
        bool f2;
        int c2 = src_cell(srcs[1], f2);
        acc = bin(KOP_ADD, acc, c2);
        for (size_t i = 2; i < srcs.size(); i++) {
            int ci = src_cell(srcs[i], f2);
            op(KOP_ADD, acc, acc, ci);
        }
        return acc;
    }
};

// ---- allocation --------------------------------------------------------------------------------------------
struct AllocResult { std::vector<uint32_t> ops; int64_t n_ops = 0; int max_slot = -1; int64_t ldg = 0, stg = 0; };

struct Allocator {
    std::vector<Cell> &cells;
    int S;                 // total scratch slots
    int first_free_slot;   // slots below are pinned
    int &n_planes;
    std::vector<int> slot_of;      // cell -> slot
    std::vector<int> cell_in;      // slot -> cell
    std::vector<char> dirty, stored, valid;   // valid: the slot holds the cell's current value
    bool pinned_valid;                         // pinned cells were computed by an earlier section
    const std::vector<double> *pool_values = nullptr;   // constants are inlined into the op stream (KOP_LDK)
    // Loads from HBM (inputs, spilled values, previous-row unknowns, source tables) are asynchronous (cp.async into the
    // slot): the allocator issues them PREFETCH_DISTANCE ops ahead of their first use and emits KOP_WAIT before that use.
    int PREFETCH_DISTANCE = 16;   // set from the scratch capacity: prefetched values must survive until they are used
    std::vector<char> pending;
    std::vector<int> pending_list;
    int64_t waits = 0;
    AllocResult out;

    Allocator(std::vector<Cell> &c, int S_, int first_free, int &planes, bool pinned_valid_)
        : cells(c), S(S_), first_free_slot(first_free), n_planes(planes), pinned_valid(pinned_valid_) {
        slot_of.assign(cells.size(), -1);
        cell_in.assign(S, -1);
        dirty.assign(cells.size(), 0);
        stored.assign(cells.size(), 0);
        valid.assign(cells.size(), 0);
        pending.assign(cells.size(), 0);
        PREFETCH_DISTANCE = std::max(4, std::min(20, (S - first_free_slot) / 3));
    }
    void emit(uint32_t opc, uint32_t d, uint32_t a, uint32_t b, uint32_t imm) {
        if (opc == KOP_LDB && pool_values) {
            // inline constant: the value travels in the op stream (next slot) instead of a table in global memory;
            // the interpreter fetches 32-op chunks, so the pair must not straddle a chunk
            if (out.n_ops % 32 == 31) { out.ops.push_back(krb_op_w0(KOP_NOP, 0, 0, 0)); out.ops.push_back(0); out.n_ops++; }
            uint64_t bits;
            const double v = (*pool_values)[imm];
            std::memcpy(&bits, &v, 8);
            out.ops.push_back(krb_op_w0(KOP_LDK, d, 0, 0));
            out.ops.push_back(0);
            out.ops.push_back((uint32_t)(bits & 0xffffffffu));
            out.ops.push_back((uint32_t)(bits >> 32));
            out.n_ops += 2;
            return;
        }
        out.ops.push_back(krb_op_w0(opc, d, a, b));
        out.ops.push_back(imm);
        out.n_ops++;
    }
    void track(int s) { out.max_slot = std::max(out.max_slot, s); }
    void mark_pending(int c) { if (!pending[c]) { pending[c] = 1; pending_list.push_back(c); } }
    void emit_wait() {
        if (pending_list.empty()) return;
        emit(KOP_WAIT, 0, 0, 0, 0);
        waits++;
        for (int c : pending_list) pending[c] = 0;
        pending_list.clear();
    }
    // issue the load that (re)materialises cell c into slot s
    bool emit_load(int c, int s, std::string &err) {
        if (!dirty[c] && cells[c].remat == R_LDB && !stored[c]) emit(KOP_LDB, s, 0, 0, cells[c].rimm);
        else if (!dirty[c] && cells[c].remat == R_LDI && !stored[c]) { emit(KOP_LDI, s, 0, 0, cells[c].rimm); mark_pending(c); }
        else if (stored[c]) { emit(KOP_LDG, s, 0, 0, (uint32_t)cells[c].home); out.ldg++; mark_pending(c); }
        else { err = "internal: read of an undefined cell"; return false; }
        return true;
    }

    bool run(const std::vector<VOp> &vops_in, std::string &err) {
        // pure loads (previous-row unknowns, source-table entries) have no operands: hoist them so that their latency
        // overlaps the ops in between
        std::vector<VOp> vops;
        vops.reserve(vops_in.size());
        for (const VOp &v : vops_in) {
            if (v.opc == KOP_LDX || v.opc == KOP_LDBS) {
                const size_t pos = vops.size() > (size_t)PREFETCH_DISTANCE ? vops.size() - PREFETCH_DISTANCE : 0;
                vops.insert(vops.begin() + pos, v);
            } else vops.push_back(v);
        }
        const int nv = (int)vops.size();
        // use lists
        std::vector<std::vector<int>> uses(cells.size());
        auto add_use = [&](int c, int i) { if (c >= 0 && (uses[c].empty() || uses[c].back() != i)) uses[c].push_back(i); };
        for (int i = 0; i < nv; i++) { add_use(vops[i].d, i); add_use(vops[i].a, i); add_use(vops[i].b, i); }
        std::vector<int> up(cells.size(), 0);   // pointer into uses
        auto next_use = [&](int c, int now) {
            auto &u = uses[c];
            int &p = up[c];
            while (p < (int)u.size() && u[p] <= now) p++;
            return p < (int)u.size() ? u[p] : INT32_MAX;
        };
        // pinned cells are resident from the start
        for (size_t c = 0; c < cells.size(); c++)
            if (cells[c].pin_slot >= 0) {
                slot_of[c] = cells[c].pin_slot; cell_in[cells[c].pin_slot] = (int)c; track(cells[c].pin_slot);
                valid[c] = pinned_valid ? 1 : 0;
            }
        // persistent home-mode cells produced by an earlier section are already in their planes
        for (size_t c = 0; c < cells.size(); c++)
            if (cells[c].persist && cells[c].home >= 0 && cells[c].pin_slot < 0) stored[c] = 1;

        for (int i = 0; i < nv; i++) {
            const VOp &v = vops[i];
            if (v.opc == VOP_PERSIST) {
                const int c = v.a;
                // a persistent value that nothing in this section consumed (e.g. an L~ entry whose pivot row has no
                // U part) still has to be materialised from its constant/input source
                if (!valid[c] && !stored[c]) {
                    int excl1[3] = {c, -1, -1};
                    int s = slot_of[c];
                    if (s < 0) s = get_slot(excl1, i, next_use, err);
                    if (s < 0) return false;
                    if (cells[c].remat == R_LDB) emit(KOP_LDB, s, 0, 0, cells[c].rimm);
                    else if (cells[c].remat == R_LDI) { emit(KOP_LDI, s, 0, 0, cells[c].rimm); mark_pending(c); }
                    else { err = "internal: persistent cell never defined"; return false; }
                    slot_of[c] = s; cell_in[s] = c; track(s); valid[c] = 1; dirty[c] = 1;
                }
                if (cells[c].pin_slot >= 0) continue;
                if (cells[c].home < 0) cells[c].home = n_planes++;
                if (slot_of[c] >= 0) {
                    if (pending[c]) emit_wait();
                    if (dirty[c] || !stored[c]) { emit(KOP_STG, 0, slot_of[c], 0, (uint32_t)cells[c].home); out.stg++; }
                    dirty[c] = 0; stored[c] = 1;
                } else if (!stored[c]) { err = "internal: persistent cell lost"; return false; }
                if (next_use(c, i) == INT32_MAX) release(c);
                continue;
            }
            // ---- lookahead: bring the HBM-resident operands of the op PREFETCH_DISTANCE ahead on chip now ----
            if (i + PREFETCH_DISTANCE < nv) {
                const VOp &f = vops[i + PREFETCH_DISTANCE];
                if (f.opc != VOP_PERSIST) {
                    int fr[3] = {f.a, f.b, f.opc == KOP_FNMA ? f.d : -1};
                    int cur_excl[3] = {v.d, v.a, v.b};
                    for (int r = 0; r < 3; r++) {
                        const int c = fr[r];
                        if (c < 0 || slot_of[c] >= 0 || cells[c].pin_slot >= 0) continue;
                        const bool loadable = (!dirty[c] && cells[c].remat == R_LDI && !stored[c]) || stored[c];
                        if (!loadable) continue;
                        const int s = get_slot(cur_excl, i, next_use, err, i + PREFETCH_DISTANCE);
                        if (s < 0) { err.clear(); continue; }          // nothing can be evicted without hurting a nearer use
                        if (!emit_load(c, s, err)) return false;
                        slot_of[c] = s; cell_in[s] = c; track(s); valid[c] = 1;
                    }
                }
            }
            const bool reads_d = (v.opc == KOP_FNMA);
            int excl[3] = {v.d, v.a, v.b};
            // operands
            int rd[3]; int nr = 0;
            if (v.a >= 0) rd[nr++] = v.a;
            if (v.b >= 0 && v.b != v.a) rd[nr++] = v.b;
            if (reads_d && v.d >= 0 && v.d != v.a && v.d != v.b) rd[nr++] = v.d;
            for (int r = 0; r < nr; r++) {
                const int c = rd[r];
                if (slot_of[c] >= 0 && valid[c]) continue;
                int s = slot_of[c];
                if (s < 0) s = get_slot(excl, i, next_use, err);
                if (s < 0) return false;
                valid[c] = 1;
                if (!emit_load(c, s, err)) return false;
                slot_of[c] = s; cell_in[s] = c; track(s);
            }
            int sd = 0;
            if (v.d >= 0) {
                if (slot_of[v.d] < 0) {
                    const int s = get_slot(excl, i, next_use, err);
                    if (s < 0) return false;
                    slot_of[v.d] = s; cell_in[s] = v.d; track(s);
                }
                sd = slot_of[v.d];
                dirty[v.d] = 1; stored[v.d] = 0; valid[v.d] = 1;
            }
            if ((v.a >= 0 && pending[v.a]) || (v.b >= 0 && pending[v.b]) || (v.d >= 0 && pending[v.d])) emit_wait();
            const int sa = v.a >= 0 ? slot_of[v.a] : 0, sb = v.b >= 0 ? slot_of[v.b] : 0;
            uint32_t bfield = (uint32_t)sb;
            if (v.opc == KOP_LDX || v.opc == KOP_XOUT) bfield = v.flags;
            emit(v.opc, (uint32_t)sd, (uint32_t)sa, bfield, v.imm);
            if ((v.opc == KOP_LDX || v.opc == KOP_LDBS) && v.d >= 0) mark_pending(v.d);
            // free dead cells
            for (int r = 0; r < 3; r++) {
                const int c = excl[r];
                if (c < 0 || slot_of[c] < 0) continue;
                if (cells[c].pin_slot >= 0) continue;
                if (next_use(c, i) == INT32_MAX) release(c);
            }
        }
        emit_wait();
        emit(KOP_END, 0, 0, 0, 0);
        return true;
    }
    void release(int c) {
        if (cells[c].pin_slot >= 0) return;
        const int s = slot_of[c];
        if (s >= 0) { cell_in[s] = -1; slot_of[c] = -1; valid[c] = 0; }
    }
    template <class NU>
    int get_slot(const int excl[3], int now, NU &next_use, std::string &err, int min_far = -1) {
        for (int s = first_free_slot; s < S; s++)
            if (cell_in[s] < 0) return s;
        int victim = -1, far = -1;
        for (int s = first_free_slot; s < S; s++) {
            const int c = cell_in[s];
            if (c == excl[0] || c == excl[1] || c == excl[2]) continue;
            if (cells[c].pin_slot >= 0) continue;
            int nu = next_use(c, now - 1 < 0 ? -1 : now - 1);
            // operands of the current op that are already resident must stay: their next use is `now`
            if (nu == now) continue;
            // prefer victims that need no store: clean remat cells or cells whose plane copy is current
            int cost_free = (!dirty[c] && (cells[c].remat != R_NONE || stored[c])) ? 1 : 0;
            int key = (nu == INT32_MAX) ? INT32_MAX : nu;
            if (key > far || (key == far && cost_free)) { far = key; victim = s; }
        }
        if (victim < 0) { err = "scratch capacity too small for one operation"; return -1; }
        if (min_far >= 0 && far <= min_far) { err = "no victim beyond the prefetch horizon"; return -1; }
        const int c = cell_in[victim];
        if (pending[c]) emit_wait();   // the slot is the target of an in-flight asynchronous load
        const int nu = next_use(c, now - 1 < 0 ? -1 : now - 1);
        const bool needed_later = (nu != INT32_MAX) || (cells[c].persist && !stored[c]);
        if (needed_later && (dirty[c] || (cells[c].remat == R_NONE && !stored[c]))) {
            if (cells[c].home < 0) cells[c].home = n_planes++;
            emit(KOP_STG, 0, (uint32_t)victim, 0, (uint32_t)cells[c].home);
            out.stg++;
            stored[c] = 1; dirty[c] = 0;
        }
        cell_in[victim] = -1; slot_of[c] = -1; valid[c] = 0;
        return victim;
    }
};

std::vector<Src> entry_srcs(const Plan &p, int entry) {
    std::vector<Src> v;
    if (entry < 0) return v;
    for (int ci : p.pattern.contribs[entry]) v.push_back(p.stamps.mat[ci].s);
    return v;
}
std::vector<Src> rhs_srcs(const Plan &p, int row) {
    std::vector<Src> v;
    for (int ri : p.pattern.rhs_contribs[row]) v.push_back(p.stamps.rhs[ri].s);
    return v;
}

// ---- block-partitioned chain tables (chain_block_kernel.cuh) ---------------------------------------------------------
// One CTA of T threads solves one lane: thread t owns pivot rows [t*M, (t+1)*M); the last row of each block is a
// separator.  Interior rows are eliminated inside the thread (left boundary carried as a fill column), the T separator
// equations form a tridiagonal system solved by parallel cyclic reduction, interiors follow by substitution.  Nothing
// but the inputs and the n results touches HBM.  Applicability is decided here, on the host:
//   * n <= T*M for one of the compiled (T, M) shapes;
//   * a row with a second super-diagonal must sit at local index <= M-3 of its block, at most one per block;
//   * the partitioned elimination, run at the nominal inputs, must agree with the sequential one to 5e-10 (scaled) —
//     the interior blocks are factorised without their left neighbour, which is only safe when they are well conditioned
//     on their own (ladders, RC lines: diagonally dominant); anything else keeps the sequential chain kernel.
typedef std::complex<double> cd;
static void chain_rows_nominal(const Plan &plan, const std::vector<uint32_t> &row_off, const std::vector<double> &nom, double w,
                               std::vector<cd> e[5]) {
    const ChainPlan &ch = plan.chain;
    const int n = plan.sym.n;
    const bool cplx = plan.is_complex;
    auto dval = [&](size_t word) { double v; std::memcpy(&v, &ch.stream[word], 8); return v; };
    for (int g = 0; g < 5; g++) e[g].assign(n, cd(0.0, 0.0));
    for (int k = 0; k < n; k++) {
        const size_t w0 = (size_t)row_off[k] * 2;
        if (ch.rowinfo[k] & 1u) {
            e[0][k] = cd(dval(w0 + 0), cplx ? w * dval(w0 + 4) : 0.0);
            e[1][k] = cd(dval(w0 + 1), cplx ? w * dval(w0 + 5) : 0.0);
            e[2][k] = cd(dval(w0 + 2), cplx ? w * dval(w0 + 6) : 0.0);
            e[4][k] = cd(dval(w0 + 3), 0.0);
        } else {
            const int nt = (int)(ch.rowinfo[k] >> 8);
            for (int t = 0; t < nt; t++) {
                const uint64_t mw = ch.stream[w0 + 2 * t];
                const uint32_t meta = (uint32_t)mw, plane = (uint32_t)(mw >> 32);
                const double K = dval(w0 + 2 * t + 1);
                cd v(0.0, 0.0);
                switch ((meta >> 4) & 0xfu) {
                case 0: v = cd(K, 0.0); break;
                case 1: v = cd(nom[plane], 0.0); break;
                case 2: v = cd(1.0 / nom[plane], 0.0); break;
                case 3: v = cd(nom[plane] / K, 0.0); break;
                case 4: v = cd(0.0, w * K); break;
                default: v = cd(0.0, w * nom[plane]); break;
                }
                if (meta & (1u << 8)) v = -v;
                e[meta & 7u][k] += v;
            }
        }
    }
}

static void build_chain_block(Plan &plan, const std::vector<uint32_t> &row_off, const std::vector<int> &input_of_elem) {
    ChainPlan &ch = plan.chain;
    ChainPlan::Block &bk = ch.block;
    bk = ChainPlan::Block();
    const int n = plan.sym.n;
    const bool cplx = plan.is_complex;
    const char *env = getenv("KRB_CHAIN_BLOCK");          // "0": never, "1": also for small systems (tests)
    if (env && env[0] == '0') { bk.why_not = "disabled by KRB_CHAIN_BLOCK=0"; return; }
    const bool force_small = env && env[0] == '1';
    if (n < 512 && !force_small) { bk.why_not = "n < 512: the sequential chain kernel is used"; return; }
    // KRB_CHAIN_BLOCK_CT256 (profiling): complex plans on 256-thread CTAs, two resident per SM (systems up to 2560 rows)
    const bool ct256 = cplx && getenv("KRB_CHAIN_BLOCK_CT256") != nullptr;
    // compiled shapes: real M = 5, 9, ... 21 (odd: the lane's result vector is contiguous AND conflict-free in shared memory),
    // complex M = 2, 4, ... 10
    const int T = cplx && !ct256 ? 512 : 256, mfirst = cplx ? 2 : 5, mstep = cplx ? 2 : 4, mmax = cplx ? 10 : 21;
    // smallest compiled M that holds the system and keeps every row with a second super-diagonal at local index <= M-3,
    // at most one per block
    int M = 0;
    for (int m = mfirst; m <= mmax && !M; m += mstep) {
        if ((int64_t)T * m < n) continue;
        bool fits = true;
        int last_block = -1;
        for (int k : ch.u2_rows) {
            if (k % m > m - 3 || k / m == last_block) { fits = false; break; }
            last_block = k / m;
        }
        if (fits) M = m;
    }
    if (!M) { bk.why_not = "no compiled (T, M) shape holds the system with its second-super-diagonal rows inside a block"; return; }
    // ---- nominal numeric check: sequential vs partitioned elimination ----
    std::vector<double> nom(std::max(plan.n_inputs, 1), 1.0);
    const double fnom = plan.opts.nominal_frequency > 0.0 ? plan.opts.nominal_frequency : 1000.0;
    for (size_t el = 0; el < input_of_elem.size(); el++) if (input_of_elem[el] >= 0) nom[input_of_elem[el]] = plan.circuit.el[el].value;
    if (plan.input_freq >= 0) nom[plan.input_freq] = fnom;
    if (plan.input_sweep >= 0 && nom[plan.input_sweep] == 0.0) nom[plan.input_sweep] = 1.0;   // a zero source would hide the right-hand side
    const double w = (2.0 * 3.14159265358979323846264338327950288) * fnom;
    std::vector<cd> e[5];
    chain_rows_nominal(plan, row_off, nom, w, e);
    const int np = T * M;
    for (int g = 0; g < 5; g++) e[g].resize(np, cd(0.0, 0.0));
    for (int k = n; k < np; k++) e[1][k] = cd(1.0, 0.0);
    const cd zero(0.0, 0.0);
    // sequential reference (what krb_chain_kernel computes)
    std::vector<cd> xs(np), u(np), u2(np), y(np);
    {
        cd pu = zero, pu2 = zero, py = zero;
        for (int k = 0; k < np; k++) {
            const cd wk = e[1][k] - e[0][k] * pu, s1 = e[2][k] - e[0][k] * pu2, yk = e[4][k] - e[0][k] * py;
            pu = s1 / wk; pu2 = e[3][k] / wk; py = yk / wk;
            u[k] = pu; u2[k] = pu2; y[k] = py;
        }
        for (int k = np - 1; k >= 0; k--) xs[k] = y[k] - (k + 1 < np ? u[k] * xs[k + 1] : zero) - (k + 2 < np ? u2[k] * xs[k + 2] : zero);
    }
    // partitioned
    std::vector<cd> xb(np), l(np), p(np), r(np), q(np);
    std::vector<cd> A(T), B(T), Cc(T), D(T);
    bool ok = true;
    auto pivot_ok = [&](cd wk, cd a, cd b) {   // wk = a - b: finite, non-zero, no catastrophic cancellation
        const double m = std::max(std::abs(a), std::abs(b));
        return std::isfinite(wk.real()) && std::isfinite(wk.imag()) && std::abs(wk) > 1e-9 * m && std::abs(wk) > 0.0;
    };
    for (int t = 0; t < T && ok; t++) {
        const int a = t * M;
        for (int j = 0; j < M - 1; j++) {
            const int k = a + j;
            cd wk, s1, yk, lk;
            if (j == 0) { wk = e[1][k]; s1 = e[2][k]; yk = e[4][k]; lk = e[0][k]; ok = ok && pivot_ok(wk, wk, zero); }
            else {
                wk = e[1][k] - e[0][k] * u[k - 1]; s1 = e[2][k] - e[0][k] * u2[k - 1]; yk = e[4][k] - e[0][k] * y[k - 1]; lk = -e[0][k] * l[k - 1];
                ok = ok && pivot_ok(wk, e[1][k], e[0][k] * u[k - 1]);
            }
            u[k] = s1 / wk; u2[k] = e[3][k] / wk; y[k] = yk / wk; l[k] = lk / wk;
        }
        const int sp = a + M - 1;
        A[t] = -e[0][sp] * l[sp - 1]; B[t] = e[1][sp] - e[0][sp] * u[sp - 1]; D[t] = e[4][sp] - e[0][sp] * y[sp - 1];
        p[sp] = zero; r[sp] = cd(-1.0, 0.0); q[sp] = zero;
        for (int j = M - 2; j >= 0; j--) {
            const int k = a + j;
            const cd p2 = j + 2 <= M - 1 ? p[k + 2] : zero, r2 = j + 2 <= M - 1 ? r[k + 2] : zero, q2 = j + 2 <= M - 1 ? q[k + 2] : zero;
            p[k] = y[k] - u[k] * p[k + 1] - u2[k] * p2;
            r[k] = -u[k] * r[k + 1] - u2[k] * r2;
            q[k] = l[k] - u[k] * q[k + 1] - u2[k] * q2;
        }
    }
    if (ok) {
        for (int t = 0; t < T; t++) {
            const cd ss = e[2][t * M + M - 1];
            const cd pn = t + 1 < T ? p[(t + 1) * M] : zero, rn = t + 1 < T ? r[(t + 1) * M] : zero, qn = t + 1 < T ? q[(t + 1) * M] : zero;
            const cd b0 = B[t];
            B[t] = b0 - ss * qn; Cc[t] = -ss * rn; D[t] = D[t] - ss * pn;
            ok = ok && pivot_ok(B[t], b0, ss * qn);
        }
        // parallel cyclic reduction, rows normalised to a unit diagonal
        std::vector<cd> a1(T), c1(T), d1(T), a2(T), c2(T), d2(T);
        for (int t = 0; t < T; t++) { a1[t] = A[t] / B[t]; c1[t] = Cc[t] / B[t]; d1[t] = D[t] / B[t]; }
        for (int s = 1; s < T && ok; s *= 2) {
            for (int t = 0; t < T; t++) {
                const cd am = t >= s ? a1[t - s] : zero, cm = t >= s ? c1[t - s] : zero, dm = t >= s ? d1[t - s] : zero;
                const cd ap = t + s < T ? a1[t + s] : zero, cp = t + s < T ? c1[t + s] : zero, dp = t + s < T ? d1[t + s] : zero;
                const cd sub_ = a1[t] * cm + c1[t] * ap;
                const cd b2 = cd(1.0, 0.0) - sub_;
                ok = ok && pivot_ok(b2, cd(1.0, 0.0), sub_);
                a2[t] = -a1[t] * am / b2; c2[t] = -c1[t] * cp / b2; d2[t] = (d1[t] - a1[t] * dm - c1[t] * dp) / b2;
            }
            a1.swap(a2); c1.swap(c2); d1.swap(d2);
        }
        for (int t = 0; t < T; t++) {
            const cd st = d1[t], sm = t > 0 ? d1[t - 1] : zero;
            for (int j = 0; j < M - 1; j++) { const int k = t * M + j; xb[k] = p[k] - r[k] * st - q[k] * sm; }
            xb[t * M + M - 1] = st;
        }
    }
    if (!ok) { bk.why_not = "partitioned elimination meets a weak pivot at the nominal inputs"; return; }
    double scale = 0.0, diff = 0.0;
    for (int k = 0; k < n; k++) scale = std::max(scale, std::abs(xs[k]));
    for (int k = 0; k < n; k++) {
        const double den = std::max(std::abs(xs[k]), 1e-6 * scale);
        if (den > 0.0) diff = std::max(diff, std::abs(xb[k] - xs[k]) / den);
        if (!std::isfinite(std::abs(xb[k]))) diff = INFINITY;
    }
    bk.nominal_diff = diff;
    if (!(diff <= 5e-10)) { bk.why_not = "partitioned and sequential elimination differ by more than 5e-10 at the nominal inputs"; return; }

    // ---- tables ----
    const int fw = cplx ? 8 : 4;                                  // words per FIXED record
    std::map<std::vector<uint64_t>, uint32_t> index;
    auto intern = [&](const std::vector<uint64_t> &rec) {
        auto it = index.find(rec);
        if (it != index.end()) return it->second;
        const uint32_t id = (uint32_t)index.size();
        index.emplace(rec, id);
        bk.dict.insert(bk.dict.end(), rec.begin(), rec.end());
        return id;
    };
    std::vector<uint64_t> pad_rec(fw, 0);
    { const double one = 1.0; std::memcpy(&pad_rec[1], &one, 8); }
    bk.rows.assign((size_t)2 * T * M, 0u);
    bk.gen_rows.clear();
    for (int k = 0; k < n; k++)
        if (!(ch.rowinfo[k] & 1u)) {
            if ((ch.rowinfo[k] >> 8) > 0x3fffu) { bk = ChainPlan::Block(); bk.why_not = "a row has more than 16383 stamp terms"; return; }
            bk.gen_rows.push_back(k);
        }
    for (int t = 0; t < T; t++)
        for (int j = 0; j < M; j++) {
            const int k = t * M + j;
            uint32_t w0, w1 = 0;
            if (k >= n) w0 = (intern(pad_rec) << 1) | 1u;
            else if (ch.rowinfo[k] & 1u) {
                std::vector<uint64_t> rec(ch.stream.begin() + (size_t)row_off[k] * 2, ch.stream.begin() + (size_t)row_off[k] * 2 + fw);
                w0 = (intern(rec) << 1) | 1u;
            } else {
                // word 1: bit 1 HAS_U2, bits 2..15 #terms, bits 16..31 slot (rank among the GENERIC rows, when it fits)
                const uint32_t slot = (uint32_t)(std::lower_bound(bk.gen_rows.begin(), bk.gen_rows.end(), k) - bk.gen_rows.begin());
                w0 = row_off[k] << 1;
                w1 = (ch.rowinfo[k] & 2u) | ((ch.rowinfo[k] >> 8) << 2) | ((slot & 0xffffu) << 16);
            }
            bk.rows[((size_t)j * T + t) * 2 + 0] = w0;
            bk.rows[((size_t)j * T + t) * 2 + 1] = w1;
        }
    bk.n_dict = (int)index.size();
    if (bk.n_dict > 0xffff) { bk = ChainPlan::Block(); bk.why_not = "dictionary index does not fit 16 bits"; return; }
    // shared-dictionary mode of the kernel: unique FIXED records + GENERIC rows fit the on-chip table
    bk.shared_dict = bk.n_dict + (int)bk.gen_rows.size() <= (cplx ? 64 : 256) && getenv("KRB_CHAIN_BLOCK_NO_SD") == nullptr;   // KRB_BLOCK_DMAX(_CPLX)
    bk.T = T; bk.M = M;
    bk.valid = true;
}

// ---- chain executor tables ------------------------------------------------------------------------------------
// Applicable when the circuit is linear, the solve is one-shot (DC / AC) and the permuted matrix has lower bandwidth
// <= 1 and upper bandwidth <= 2:  Lpat[k] in {k-1},  Upat[k] in {k+1, k+2}.
void build_chain_plan(Plan &plan, const std::vector<int> &input_of_elem) {
    ChainPlan &ch = plan.chain;
    ch = ChainPlan();
    const Symbolic &S = plan.sym;
    const int n = S.n;
    if (plan.circuit.nonlinear || plan.mode == KRB_MODE_TRAN || n == 0) return;
    for (int k = 0; k < n; k++) {
        for (int j : S.Lpat[k]) if (j != k - 1) return;
        for (int l : S.Upat[k]) if (l != k + 1 && l != k + 2) return;
    }
    auto emit_terms = [&](const std::vector<Src> &srcs, uint32_t target, std::vector<ChainTerm> &out) -> bool {
        bool have_const = false;     // a constant run being folded (consecutive lane-invariant contributions)
        double acc = 0.0;
        auto flush = [&]() { if (have_const) { out.push_back({target, 0u, acc}); have_const = false; acc = 0.0; } };
        for (const Src &s : srcs) {
            const krb_element &el = plan.circuit.el[s.elem];
            const int pl = input_of_elem[s.elem];
            double cv = 0.0;
            bool is_const = true;
            uint32_t kind = 0;
            double K = 0.0;
            switch (s.kind) {
            case S_ONE: cv = 1.0; break;
            case S_G: if (pl < 0) cv = 1.0 / el.value; else { is_const = false; kind = 2; } break;
            case S_VAL: if (pl < 0) cv = el.value; else { is_const = false; kind = 1; } break;
            case S_VOH: if (pl < 0) cv = el.value / plan.opts.time_step; else { is_const = false; kind = 3; K = plan.opts.time_step; } break;
            case S_ACAMP: cv = el.ac_amplitude; break;
            case S_JWV: is_const = false; if (pl < 0) { kind = 4; K = el.value; } else kind = 5; break;
            default: return false;   // nonlinear / history sources never reach the chain executor
            }
            if (is_const) {
                if (s.neg) cv = -cv;
                acc = have_const ? acc + cv : 0.0 + cv;   // same left-to-right order as sum_triplets
                have_const = true;
            } else {
                flush();   // keep the summation order: the running constant goes first
                out.push_back({target | (kind << 4) | (s.neg ? (1u << 8) : 0u), (uint32_t)(pl < 0 ? 0 : pl), K});
            }
        }
        flush();
        return true;
    };
    auto dbits = [](double v) { uint64_t b; std::memcpy(&b, &v, 8); return b; };
    const bool cplx = plan.is_complex;
    int n_u2 = 0, n_fixed = 0;
    std::vector<uint32_t> row_off(n);                             // 16-byte unit offset of each row's record / term run
    for (int k = 0; k < n; k++) {
        row_off[k] = (uint32_t)(ch.stream.size() / 2);
        std::vector<ChainTerm> row;
        if (!S.Lpat[k].empty() && !emit_terms(entry_srcs(plan, S.Lent[k][0]), 0u, row)) return;
        if (!emit_terms(entry_srcs(plan, S.Dent[k]), 1u, row)) return;
        bool has_u2 = false;
        for (size_t t = 0; t < S.Upat[k].size(); t++) {
            const bool far = S.Upat[k][t] == k + 2;
            has_u2 = has_u2 || far;
            if (!emit_terms(entry_srcs(plan, S.Uent[k][t]), far ? 3u : 2u, row)) return;
        }
        if (!emit_terms(rhs_srcs(plan, S.row_order[k]), 4u, row)) return;
        // FIXED record: per entry one folded constant run and at most one j*w*K term (components are independent, so the
        // reference's summation order is preserved); no second super-diagonal, no per-lane inputs.
        bool fixed = !has_u2;
        double re[5] = {0, 0, 0, 0, 0}, jw[5] = {0, 0, 0, 0, 0};
        int nre[5] = {0, 0, 0, 0, 0}, njw[5] = {0, 0, 0, 0, 0};
        for (const ChainTerm &t : row) {
            const uint32_t tgt = t.meta & 7u, kind = (t.meta >> 4) & 0xfu;
            if (kind == 0) { re[tgt] = t.K; nre[tgt]++; }
            else if (kind == 4 && cplx) { jw[tgt] = (t.meta & (1u << 8)) ? -t.K : t.K; njw[tgt]++; }
            else fixed = false;
        }
        for (int g = 0; g < 5; g++) if (nre[g] > 1 || njw[g] > 1) fixed = false;
        if (njw[4] || nre[3] || njw[3]) fixed = false;
        if (fixed) {
            ch.rowinfo.push_back(1u);
            ch.stream.push_back(dbits(re[0])); ch.stream.push_back(dbits(re[1]));
            ch.stream.push_back(dbits(re[2])); ch.stream.push_back(dbits(re[4]));
            if (cplx) {
                ch.stream.push_back(dbits(jw[0])); ch.stream.push_back(dbits(jw[1]));
                ch.stream.push_back(dbits(jw[2])); ch.stream.push_back(dbits(0.0));
            }
            n_fixed++;
        } else {
            ch.rowinfo.push_back((has_u2 ? 2u : 0u) | ((uint32_t)row.size() << 8));
            for (const ChainTerm &t : row) {
                ch.stream.push_back((uint64_t)t.meta | ((uint64_t)t.plane << 32));
                ch.stream.push_back(dbits(t.K));
            }
            if (has_u2) { n_u2++; ch.u2_rows.push_back(k); }
        }
    }
    for (int k = 0; k < n; k++) {                                 // run table
        const uint32_t fx = ch.rowinfo[k] & 1u;
        if (!ch.runs.empty() && (ch.runs.back() & 1u) == fx) ch.runs.back() += 2u;
        else ch.runs.push_back(2u | fx);
    }
    ch.rowinfo.push_back(1u);                                   // sentinel row (prefetched, never used)
    for (int i = 0; i < 2 * 1100; i++) ch.stream.push_back(0);   // the kernel reads one record and prefetches up to 1024 units past the end
    for (int i = 0; i < 1100; i++) ch.rowinfo.push_back(1u);
    ch.unknown_of_row = S.col_order;
    ch.identity_order = true;
    for (int k = 0; k < n; k++) if (S.col_order[k] != k) ch.identity_order = false;
    ch.n_u2 = n_u2;
    ch.n_fixed = n_fixed;
    ch.valid = true;
    build_chain_block(plan, row_off, input_of_elem);
}

// ---- numeric check of the static pivot plan of a lane program ------------------------------------------------------------------
// The row matching is a maximum-product transversal of |values| and the elimination order comes from the pattern alone, so nothing
// above sees SIGNS: a floating voltage source whose terminals are also joined through companion conductances (V between nodes a, b;
// C from a to c; R from c to b) gives a pivot sequence in which the branch current's pivot cancels EXACTLY, for any element values
// (the source's KCL rows sum to an equation without the current) — the reference never notices because faer pivots numerically at
// every factorisation (op.rs:50).  So the up-looking elimination of the lane program is run once on the host at the nominal values,
// and a pivot that cancels to rounding level (|d| <= 1e-13 of the largest term of its own sum), vanishes or overflows rejects the plan.
static bool static_pivots_ok(const Symbolic &S, const std::vector<cd> &val) {
    const int n = S.n;
    std::vector<cd> W(n, cd(0.0, 0.0));
    std::vector<std::vector<cd>> Uhat(n);
    for (int k = 0; k < n; k++) {
        const auto &Lp = S.Lpat[k];
        const auto &Up = S.Upat[k];
        for (size_t t = 0; t < Lp.size(); t++) W[Lp[t]] = S.Lent[k][t] >= 0 ? val[S.Lent[k][t]] : cd(0.0, 0.0);
        for (size_t t = 0; t < Up.size(); t++) W[Up[t]] = S.Uent[k][t] >= 0 ? val[S.Uent[k][t]] : cd(0.0, 0.0);
        W[k] = S.Dent[k] >= 0 ? val[S.Dent[k]] : cd(0.0, 0.0);
        double big = std::abs(W[k]);
        for (int j : Lp) {                                   // ascending: row j of U^ is final
            const cd l = W[j];
            const auto &Uj = S.Upat[j];
            for (size_t t = 0; t < Uj.size(); t++) {
                const cd term = l * Uhat[j][t];
                if (Uj[t] == k) big = std::max(big, std::abs(term));
                W[Uj[t]] -= term;
            }
        }
        const double d = std::abs(W[k]);
        if (!(d > 0.0) || !std::isfinite(d) || !(d > 1e-13 * big)) return false;
        Uhat[k].resize(Up.size());
        for (size_t t = 0; t < Up.size(); t++) Uhat[k][t] = W[Up[t]] / W[k];
    }
    return true;
}

// Repair: what a numerically pivoting LU would choose at the nominal values — dense elimination in the given unknown order with
// threshold partial pivoting over the equations not used yet; the equation the transversal matched to the unknown is kept whenever it
// is within a factor 10 of the column maximum (the plan stays as close to the sparse one as the numbers allow).
static bool repivot_at_nominal(const Pattern &p, const std::vector<cd> &val, const std::vector<int> &order, std::vector<int> &match_row) {
    const int n = p.n;
    if (n > 2048) return false;                              // dense n x n work matrix: lane programs are far smaller
    std::vector<cd> A((size_t)n * n, cd(0.0, 0.0));          // A[row * n + position of the unknown in `order`]
    std::vector<int> pos_of_col(n);
    for (int k = 0; k < n; k++) pos_of_col[order[k]] = k;
    for (int col = 0; col < n; col++)
        for (int q = p.colptr[col]; q < p.colptr[col + 1]; q++) A[(size_t)p.rowidx[q] * n + pos_of_col[col]] = val[q];
    std::vector<char> used(n, 0);
    std::vector<int> chosen(n, -1);
    for (int k = 0; k < n; k++) {
        double best = 0.0;
        int rbest = -1;
        for (int r = 0; r < n; r++)
            if (!used[r]) { const double a = std::abs(A[(size_t)r * n + k]); if (a > best) { best = a; rbest = r; } }
        if (rbest < 0 || !(best > 0.0) || !std::isfinite(best)) return false;     // numerically singular at the nominal point
        const int pref = match_row[order[k]];
        if (!used[pref] && std::abs(A[(size_t)pref * n + k]) >= 0.1 * best) rbest = pref;
        used[rbest] = 1;
        chosen[k] = rbest;
        const cd piv = A[(size_t)rbest * n + k];
        for (int r = 0; r < n; r++) {
            if (used[r]) continue;
            const cd f = A[(size_t)r * n + k] / piv;
            if (f == cd(0.0, 0.0)) continue;
            for (int l = k + 1; l < n; l++) A[(size_t)r * n + l] -= f * A[(size_t)rbest * n + l];
        }
    }
    for (int k = 0; k < n; k++) match_row[order[k]] = chosen[k];
    return true;
}

}  // namespace

int analyze(const Circuit &c, int mode, const krb_plan_opts &opts_in, Plan &plan, std::string &err) {
    plan.circuit = c;
    plan.mode = mode;
    plan.opts = opts_in;
    plan.is_complex = (mode == KRB_MODE_AC);
    plan.varied.assign(opts_in.varied_elements, opts_in.varied_elements + (opts_in.n_varied > 0 ? opts_in.n_varied : 0));
    plan.opts.varied_elements = nullptr;
    const int n = c.n;
    if (mode == KRB_MODE_TRAN && !(plan.opts.time_step > 0.0)) { err = "transient plan needs time_step > 0"; return KRB_ERR_INVALID_ARGUMENT; }
    if (plan.opts.integrator < KRB_INTEGRATOR_BE_KRETS || plan.opts.integrator > KRB_INTEGRATOR_BE) { err = "unknown integrator"; return KRB_ERR_INVALID_ARGUMENT; }
    const bool trap = (mode == KRB_MODE_TRAN && plan.opts.integrator == KRB_INTEGRATOR_TRAP);
    plan.h_comp = trap ? plan.opts.time_step / 2.0 : plan.opts.time_step;
    plan.state_elems.clear();
    if (trap) for (int e = 0; e < (int)c.el.size(); e++) if (c.el[e].kind == KRB_C) plan.state_elems.push_back(e);
    for (int e : plan.varied) {
        if (e < 0 || e >= (int)c.el.size()) { err = "varied element ordinal out of range"; return KRB_ERR_INVALID_ARGUMENT; }
        const int k = c.el[e].kind;
        if (!(k == KRB_R || k == KRB_C || k == KRB_L || k == KRB_V || k == KRB_I)) { err = "only R/C/L/V/I values can vary per lane"; return KRB_ERR_INVALID_ARGUMENT; }
    }
    int rc = build_stamps(c, mode, plan.opts.integrator, plan.stamps);
    if (rc) { err = "element not supported in this analysis mode (the reference panics here)"; return rc; }
    // A voltage source without a waveform has transient value = its DC value (voltage_source.rs:244-252), so when that value varies per
    // lane (Monte-Carlo) the transient right-hand side must read the lane's input plane like the DC one does — the per-row source table
    // (S_WAVE) is lane-invariant and would keep the nominal value on every lane.  Sources with a PULSE / SIN waveform stay table-driven
    // (their DC value only enters the operating point of row 0).
    if (mode == KRB_MODE_TRAN)
        for (RhsContrib &rc : plan.stamps.rhs)
            if (rc.s.kind == S_WAVE && c.el[rc.s.elem].waveform == KRB_WAVE_NONE &&
                std::find(plan.varied.begin(), plan.varied.end(), rc.s.elem) != plan.varied.end()) rc.s.kind = S_VAL;
    build_pattern(n, plan.stamps, plan.pattern);

    // ---- static pivot plan ----
    std::vector<double> mag;
    const double fnom = plan.opts.nominal_frequency > 0.0 ? plan.opts.nominal_frequency : 1000.0;
    nominal_magnitudes(c, plan.stamps, plan.pattern, mode, plan.h_comp, fnom, mag);
    std::vector<int> match_row, order;
    if (!weighted_matching(plan.pattern, mag, match_row)) { err = "matrix is structurally singular"; return KRB_ERR_DECOMPOSITION_FAILED; }
    // Band-friendly candidates first (netlist order, reverse Cuthill-McKee both ways): ladders and lines come out with
    // lower bandwidth 1 / upper bandwidth 2 and no fill, which the native chain executor needs; otherwise minimum degree.
    bool banded = false;
    {
        std::vector<int> cand(n);
        for (int i = 0; i < n; i++) cand[i] = i;
        if (band_ok(plan.pattern, match_row, cand)) { order = cand; banded = true; }
        if (!banded && n <= 2000000) {
            rcm_order(plan.pattern, match_row, cand);
            if (band_ok(plan.pattern, match_row, cand)) { order = cand; banded = true; }
            else {
                std::reverse(cand.begin(), cand.end());
                if (band_ok(plan.pattern, match_row, cand)) { order = cand; banded = true; }
            }
        }
        if (!banded) minimum_degree(plan.pattern, match_row, order);
    }
    // ---- executor selection: the one-thread-per-lane executors unroll the whole elimination into a lane program, which is
    // right for ladders, filters and small networks; meshes go to the supernodal front executor (front_kernels.cuh) ----
    bool use_front = (plan.opts.exec == KRB_EXEC_STAGED);
    if (use_front && c.nonlinear) { err = "the front executor (KRB_EXEC_STAGED) handles linear circuits only"; return KRB_ERR_UNSUPPORTED; }
    if (!use_front) {
        symbolic_lu(plan.pattern, match_row, order, plan.sym);
        // numeric check of the static pivots at the nominal values; a cancelling sequence is re-pivoted (see static_pivots_ok).  Plans
        // that go to the front executor below are skipped: the fronts pivot numerically per lane, and their pivot-free symbolic LU is large
        if (!(plan.opts.exec == KRB_EXEC_FUSED && !banded && !c.nonlinear && plan.sym.flops > 250000)) {
            std::vector<cd> val;
            nominal_values(c, plan.stamps, plan.pattern, mode, plan.h_comp, fnom, val);
            const bool no_check = getenv("KRB_NO_PIVOT_CHECK") != nullptr;             // tests: the plan as the transversal alone gives it
            if (!no_check && !static_pivots_ok(plan.sym, val)) {
                std::vector<int> order2 = order, match2 = match_row;
                if (banded) minimum_degree(plan.pattern, match2, order2);               // the band orders have no freedom left
                if (repivot_at_nominal(plan.pattern, val, order2, match2)) {
                    Symbolic s2;
                    symbolic_lu(plan.pattern, match2, order2, s2);
                    if (static_pivots_ok(s2, val)) { match_row = match2; order = order2; plan.sym = s2; banded = false; plan.repivoted = true; }
                }
                if (getenv("KRB_PIVOT_DEBUG")) fprintf(stderr, "krets_b200: static pivot sequence cancels at the nominal values (mode %d, n = %d): %s\n", mode, n, plan.repivoted ? "re-pivoted" : "no repair found");
            }
        }
        if (plan.opts.exec == KRB_EXEC_FUSED && !banded && !c.nonlinear && plan.sym.flops > 250000) use_front = true;
    }
    if (use_front) {
        // the fronts pivot dynamically, so prefer diagonal static pivots (structural symmetry => far less fill), then reorder
        double bonus = 6.9;   // ln(1e3): growth per pivot stays <= 1e3 at the nominal point (AC: plan at the lowest frequency of the sweep)
        if (const char *e = getenv("KRB_FRONT_DIAG_BONUS")) bonus = atof(e);
        if (!weighted_matching(plan.pattern, mag, match_row, bonus)) { err = "matrix is structurally singular"; return KRB_ERR_DECOMPOSITION_FAILED; }
        minimum_degree(plan.pattern, match_row, order);
        // DC: inductors are 0 V branches whose own equation has no diagonal (stampable.rs:688-701), like V sources (:972-980).  A
        // row matching makes the pattern unsymmetric and the symmetrised fill grows 2.7x on a mesh.  Instead keep every row in
        // place and eliminate each such branch current right after one of its terminal nodes (2x2-pivot compression of KKT-like
        // systems): the node's elimination fills the missing diagonal, and the fronts pivot dynamically should it come out small.
        if (mode == KRB_MODE_DC && getenv("KRB_FRONT_NO_PAIRING") == nullptr) {
            const Pattern &pt = plan.pattern;
            std::vector<char> has_diag(n, 0), taken(n, 0);
            auto has_entry = [&](int row, int col) {
                auto b = pt.rowidx.begin() + pt.colptr[col], e = pt.rowidx.begin() + pt.colptr[col + 1];
                return std::binary_search(b, e, row);
            };
            for (int u = 0; u < n; u++) has_diag[u] = has_entry(u, u) && mag[(int)(std::lower_bound(pt.rowidx.begin() + pt.colptr[u], pt.rowidx.begin() + pt.colptr[u + 1], u) - pt.rowidx.begin())] > 0.0;
            std::vector<int> leader_of(n, -1);
            bool ok = true, any = false;
            for (int u = 0; u < n && ok; u++) {
                if (has_diag[u]) continue;
                any = true;
                int pick = -1;
                for (int q = pt.colptr[u]; q < pt.colptr[u + 1] && pick < 0; q++) {
                    const int v = pt.rowidx[q];               // A[v, u] != 0
                    if (v != u && has_diag[v] && !taken[v] && has_entry(u, v)) pick = v;
                }
                if (pick < 0) ok = false;
                else { leader_of[u] = pick; taken[pick] = 1; }
            }
            if (ok && any) {
                for (int j = 0; j < n; j++) match_row[j] = j;
                minimum_degree_paired(plan.pattern, leader_of, order);
            }
        }
    }
    const Symbolic &S = plan.sym;

    // ---- per-lane inputs ----
    Builder B(plan);
    B.input_of_elem.assign(c.el.size(), -1);
    plan.n_inputs = 0;
    plan.input_sweep = plan.input_freq = -1;
    if (mode == KRB_MODE_AC) plan.input_freq = plan.n_inputs++;
    if (mode == KRB_MODE_DC && plan.opts.sweep_element >= 0) {
        if (plan.opts.sweep_element >= (int)c.el.size()) { err = "sweep element out of range"; return KRB_ERR_ELEMENT_NOT_FOUND; }
        const int k = c.el[plan.opts.sweep_element].kind;
        if (k != KRB_V && k != KRB_I) { err = "DC sweep element must be a voltage or current source"; return KRB_ERR_INVALID_ELEMENT_FORMAT; }
        plan.input_sweep = plan.n_inputs++;
        B.input_of_elem[plan.opts.sweep_element] = plan.input_sweep;
    }
    plan.input_of_varied.clear();
    for (int e : plan.varied) {
        if (B.input_of_elem[e] >= 0) { plan.input_of_varied.push_back(B.input_of_elem[e]); continue; }
        B.input_of_elem[e] = plan.n_inputs++;
        plan.input_of_varied.push_back(B.input_of_elem[e]);
    }

    if (use_front) {
        build_front_plan(plan.pattern, match_row, order, plan.front);
        const FrontPlan &fp = plan.front;
        plan.sym = Symbolic();
        plan.sym.n = n;
        plan.sym.row_order = fp.row_order;
        plan.sym.col_order = fp.col_order;
        plan.sym.nnz_lu = fp.nnz_lu;
        plan.sym.flops = fp.flops;
        plan.front_input_of_elem = B.input_of_elem;
        plan.wave_elems.clear();
        for (const RhsContrib &rc : plan.stamps.rhs)
            if (rc.s.kind == S_WAVE && std::find(plan.wave_elems.begin(), plan.wave_elems.end(), rc.s.elem) == plan.wave_elems.end()) plan.wave_elems.push_back(rc.s.elem);
        plan.use_front = true;
        plan.use_chain = false;
        plan.has_step_sections = false;
        const int64_t sz = plan.is_complex ? 16 : 8;
        plan.flops_per_point = fp.flops * (plan.is_complex ? 4 : 1);
        plan.alg_bytes_per_point = 8 * (int64_t)std::max(1, plan.n_inputs) + sz * (4 * fp.nnz_lu + 3 * (int64_t)n);
        // what the kernels move per point: stamp A out + in, factor blocks out, contribution blocks out + in, solve reads the factor once
        plan.hbm_bytes_per_point = 8 * (int64_t)std::max(1, plan.n_inputs) + sz * (2 * (int64_t)plan.pattern.nnz + 2 * fp.lu_size + 2 * fp.cb_size + 4 * (int64_t)n);
        plan.vm_hbm_bytes_per_point = 0;
        return KRB_OK;
    }

    // ---- emission ----
    const bool split = (mode == KRB_MODE_TRAN) && !c.nonlinear;   // linear transient: factor once, then one solve per step
    plan.has_step_sections = split;
    const int which_guess = (mode == KRB_MODE_AC) ? 0 : 1;
    Section main_sec(B, which_guess), fac_sec(B, which_guess), step_sec(B, which_guess);
    std::vector<std::vector<int>> Ucell(n), Lcell(n);
    std::vector<int> yhat(n, -1), invd(n, -1);

    auto emit_factor_rows = [&](Section &sec, bool with_rhs, bool persist) {
        std::vector<int> W(n, -1);
        for (int k = 0; k < n; k++) {
            const auto &Lp = S.Lpat[k];
            const auto &Up = S.Upat[k];
            for (size_t t = 0; t < Lp.size(); t++) W[Lp[t]] = sec.sum_cell(entry_srcs(plan, S.Lent[k][t]));
            W[k] = sec.sum_cell(entry_srcs(plan, S.Dent[k]));
            for (size_t t = 0; t < Up.size(); t++) W[Up[t]] = sec.sum_cell(entry_srcs(plan, S.Uent[k][t]));
            int Y = -1;
            if (with_rhs) Y = sec.sum_cell(rhs_srcs(plan, S.row_order[k]));
            for (size_t t = 0; t < Lp.size(); t++) {
                const int j = Lp[t];
                const int wj = W[j];
                for (size_t q = 0; q < S.Upat[j].size(); q++) sec.op(KOP_FNMA, W[S.Upat[j][q]], wj, Ucell[j][q]);
                if (with_rhs) sec.op(KOP_FNMA, Y, wj, yhat[j]);
                if (persist) { B.cells[wj].persist = true; sec.vops.push_back({VOP_PERSIST, 0, -1, wj, -1, 0}); }
            }
            Lcell[k].clear();
            for (size_t t = 0; t < Lp.size(); t++) Lcell[k].push_back(W[Lp[t]]);
            const int inv = sec.op(KOP_RCP, B.new_cell(), W[k]);
            invd[k] = inv;
            Ucell[k].clear();
            for (size_t t = 0; t < Up.size(); t++) {
                const int u = W[Up[t]];
                sec.op(KOP_MULG, u, u, inv);        // u^ = u / pivot, growth-tracked (KRB_LANE_INACCURATE)
                Ucell[k].push_back(u);
                if (persist) { B.cells[u].persist = true; sec.vops.push_back({VOP_PERSIST, 0, -1, u, -1, 0}); }
            }
            if (with_rhs) { sec.op(KOP_MUL, Y, Y, inv); yhat[k] = Y; }
            if (persist) { B.cells[inv].persist = true; sec.vops.push_back({VOP_PERSIST, 0, -1, inv, -1, 0}); }
        }
    };
    auto emit_back = [&](Section &sec, std::map<int, int> *carry_out = nullptr) {
        std::vector<int> X(n, -1);
        std::vector<std::pair<int, int>> carry_updates;   // (persistent cell, new value cell)
        for (int k = n - 1; k >= 0; k--) {
            const int x = yhat[k];
            for (size_t t = 0; t < S.Upat[k].size(); t++) sec.op(KOP_FNMA, x, Ucell[k][t], X[S.Upat[k][t]]);
            const int u = S.col_order[k];
            sec.op(KOP_XOUT, -1, x, -1, (uint32_t)u, c.is_current[u] ? 1 : 0);
            X[k] = x;
        }
        // trapezoidal rule: carry each capacitor's current i_n = (2C/h)((v+ - v-)_n - (v+ - v-)_{n-1}) - i_{n-1} in a state column
        if (!plan.state_elems.empty()) {
            std::vector<int> pos_of_unknown(n);
            for (int k = 0; k < n; k++) pos_of_unknown[S.col_order[k]] = k;
            for (size_t j = 0; j < plan.state_elems.size(); j++) {
                const int e = plan.state_elems[j];
                const krb_element &el = c.el[e];
                const int g = sec.varied(e) ? sec.cached(Section::T_VOH, e, false, &Section::make_voh) : sec.make_voh(e);
                int vn, vp;
                if (el.ip >= 0 && el.im >= 0) { vn = sec.bin(KOP_SUB, X[pos_of_unknown[el.ip]], X[pos_of_unknown[el.im]]); vp = sec.bin(KOP_SUB, sec.ldx(0, el.ip, false), sec.ldx(0, el.im, false)); }
                else if (el.ip >= 0) { vn = X[pos_of_unknown[el.ip]]; vp = sec.ldx(0, el.ip, false); }
                else if (el.im >= 0) { vn = sec.un(KOP_NEG, X[pos_of_unknown[el.im]]); vp = sec.un(KOP_NEG, sec.ldx(0, el.im, false)); }
                else continue;
                const int dv = sec.bin(KOP_SUB, vn, vp);
                const int gi = sec.bin(KOP_MUL, g, dv);
                const int inew = sec.bin(KOP_SUB, gi, sec.ldx(0, n + (int)j, false));
                sec.op(KOP_XOUT, -1, inew, -1, (uint32_t)(n + j), 4);   // flag 4: state column (no convergence test, not delivered)
                if (carry_out && carry_out->count(n + (int)j)) carry_updates.push_back({(*carry_out)[n + (int)j], inew});
            }
        }
        if (carry_out) {
            // the step's reads of the carried previous-row values are all above; now roll them forward
            for (int k = 0; k < n; k++) {
                auto it = carry_out->find(S.col_order[k]);
                if (it != carry_out->end()) carry_updates.push_back({it->second, X[k]});
            }
            for (auto &cu : carry_updates) sec.op(KOP_MOV, cu.first, cu.second);
        }
    };

    if (!split) {
        emit_factor_rows(main_sec, true, false);
        emit_back(main_sec, nullptr);
    } else {
        // Carry mode: when factors + state fit the scratch, the per-step section touches HBM only to deliver results.
        int Scap0 = plan.opts.scratch_slots > 0 ? plan.opts.scratch_slots : (plan.is_complex ? 28 : 48);
        Scap0 = std::max(8, std::min(Scap0, KRB_MAX_SLOTS));
        const int width = n + (int)plan.state_elems.size();
        const bool carry_mode = (S.nnz_lu + width + (int64_t)plan.varied.size() + 12 <= Scap0);
        std::map<int, int> carry, coef_voh;
        emit_factor_rows(fac_sec, false, true);
        if (carry_mode) {
            // which previous-row columns does a step read?  (capacitor / inductor history, trapezoidal state)
            std::vector<char> need(width, 0);
            for (const RhsContrib &rc : plan.stamps.rhs) {
                const krb_element &el = c.el[rc.s.elem];
                if (rc.s.kind == S_CAP_HIST || rc.s.kind == S_IND_VPREV) { if (el.ip >= 0) need[el.ip] = 1; if (el.im >= 0) need[el.im] = 1; }
                if (rc.s.kind == S_IND_HIST && el.ic >= 0) need[el.ic] = 1;
            }
            for (size_t j = 0; j < plan.state_elems.size(); j++) {
                need[n + j] = 1;
                const krb_element &el = c.el[plan.state_elems[j]];
                if (el.ip >= 0) need[el.ip] = 1;
                if (el.im >= 0) need[el.im] = 1;
            }
            for (int u = 0; u < width; u++)
                if (need[u]) {
                    const int cell = fac_sec.ldx(0, u, false);            // first step: previous row = the operating point (row 0)
                    B.cells[cell].persist = true;
                    fac_sec.vops.push_back({VOP_PERSIST, 0, -1, cell, -1, 0});
                    carry[u] = cell;
                }
            for (int e : plan.varied) {
                const int k = c.el[e].kind;
                if (k != KRB_C && k != KRB_L) continue;
                const int cell = fac_sec.cached(Section::T_VOH, e, false, &Section::make_voh);
                if (!B.cells[cell].persist) { B.cells[cell].persist = true; fac_sec.vops.push_back({VOP_PERSIST, 0, -1, cell, -1, 0}); }
                coef_voh[e] = cell;
            }
            step_sec.carry = &carry;
            step_sec.coef_voh = &coef_voh;
        }
        for (int k = 0; k < n; k++) {
            int Y = step_sec.sum_cell(rhs_srcs(plan, S.row_order[k]));
            for (size_t t = 0; t < S.Lpat[k].size(); t++) step_sec.op(KOP_FNMA, Y, Lcell[k][t], yhat[S.Lpat[k][t]]);
            step_sec.op(KOP_MUL, Y, Y, invd[k]);
            yhat[k] = Y;
        }
        emit_back(step_sec, carry_mode ? &carry : nullptr);
    }

    // ---- allocation ----
    int Scap = plan.opts.scratch_slots > 0 ? plan.opts.scratch_slots : (plan.is_complex ? 28 : 48);
    Scap = std::max(8, std::min(Scap, KRB_MAX_SLOTS));
    plan.n_planes = 0;
    int first_free = 0;
    if (split) {
        int64_t npersist = 0;
        for (auto &cl : B.cells) npersist += cl.persist ? 1 : 0;
        if (npersist + 10 <= Scap) {
            int s = 0;
            for (auto &cl : B.cells) if (cl.persist) cl.pin_slot = s++;
            first_free = s;
        }
    }
    int max_slot = -1;
    std::vector<double> pool_values;
    evaluate_pool(plan, plan.circuit, plan.h_comp, pool_values);
    auto run_alloc = [&](Section &sec, Program &prog, bool pinned_valid) -> bool {
        Allocator A(B.cells, Scap, first_free, plan.n_planes, pinned_valid);
        A.pool_values = &pool_values;
        if (!A.run(sec.vops, err)) return false;
        prog.ops.swap(A.out.ops);
        prog.n_ops = A.out.n_ops;
        max_slot = std::max(max_slot, A.out.max_slot);
        plan.hbm_bytes_per_point += (A.out.ldg + A.out.stg);   // counted in values here, scaled below
        return true;
    };
    plan.hbm_bytes_per_point = 0;
    if (!split) {
        if (!run_alloc(main_sec, plan.sec[0], false)) return KRB_ERR_INTERNAL;
    } else {
        if (!run_alloc(fac_sec, plan.sec[1], false)) return KRB_ERR_INTERNAL;
        if (!run_alloc(step_sec, plan.sec[2], true)) return KRB_ERR_INTERNAL;
    }
    plan.n_slots = max_slot + 1;
    build_chain_plan(plan, B.input_of_elem);
    const int64_t s = plan.is_complex ? 16 : 8;
    plan.hbm_bytes_per_point = plan.hbm_bytes_per_point * s + (int64_t)n * s + 8 * (int64_t)plan.n_inputs;
    plan.flops_per_point = S.flops * (plan.is_complex ? 4 : 1);
    plan.alg_bytes_per_point = 8 * (int64_t)std::max(1, plan.n_inputs) + s * (4 * S.nnz_lu + 3 * (int64_t)n);
    plan.vm_hbm_bytes_per_point = plan.hbm_bytes_per_point;
    plan.use_chain = plan.chain.valid && plan.opts.exec != KRB_EXEC_LANE_VM;
    // small linear transient plans are stepped by the native dense kernel (dense_tran_kernel.cuh) instead of the interpreter
    plan.dense = DensePlan();
    if (split && plan.opts.exec == KRB_EXEC_FUSED && n >= 1 && n <= 8 && plan.state_elems.size() <= 8 && plan.wave_elems.size() <= 4) {
        DensePlan &dp = plan.dense;
        std::vector<int> pos_of_col(n), pos_of_row(n);
        for (int k = 0; k < n; k++) { pos_of_col[S.col_order[k]] = k; pos_of_row[S.row_order[k]] = k; }
        auto pos = [&](int u) { return u >= 0 ? pos_of_col[u] : -1; };
        dp.elems.resize(c.el.size());
        for (size_t e = 0; e < c.el.size(); e++) {
            DensePlan::Elem &de = dp.elems[e];
            de.value = c.el[e].value; de.in_plane = B.input_of_elem[e];
            de.pp = pos(c.el[e].ip); de.pm = pos(c.el[e].im); de.pc = pos(c.el[e].ic);
            de.wave_col = -1; de.state_idx = -1;
        }
        for (size_t k = 0; k < plan.wave_elems.size(); k++) dp.elems[plan.wave_elems[k]].wave_col = (int)k;
        for (size_t j = 0; j < plan.state_elems.size(); j++) dp.elems[plan.state_elems[j]].state_idx = (int)j;
        bool ok = true;
        for (const MatContrib &mc : plan.stamps.mat) {
            if (mc.s.kind != S_ONE && mc.s.kind != S_G && mc.s.kind != S_VAL && mc.s.kind != S_VOH) ok = false;
            dp.mat_terms.push_back({pos_of_row[mc.row], pos_of_col[mc.col], (int)mc.s.kind | (mc.s.neg ? 0x100 : 0), mc.s.elem});
        }
        for (const RhsContrib &rc : plan.stamps.rhs) {
            const SrcKind kd = rc.s.kind;
            if (kd != S_VAL && kd != S_WAVE && kd != S_CAP_HIST && kd != S_IND_HIST && kd != S_CAP_IPREV && kd != S_IND_VPREV) ok = false;
            dp.rhs_terms.push_back({pos_of_row[rc.row], 0, (int)kd | (rc.s.neg ? 0x100 : 0), rc.s.elem});
        }
        dp.valid = ok;
    }
    if (plan.use_chain) plan.hbm_bytes_per_point = s * (5 * (int64_t)n + 2 * plan.chain.n_u2) + 8 * (int64_t)plan.n_inputs;   // 2n(+) out, back, n results
    if (plan.use_chain && plan.chain.block.valid) plan.hbm_bytes_per_point = s * (int64_t)n + 8 * (int64_t)plan.n_inputs;   // block kernel: inputs in, n results out
    return KRB_OK;
}

}  // namespace krb
