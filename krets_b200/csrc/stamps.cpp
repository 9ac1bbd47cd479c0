// Symbolic MNA stamps: which (row, col) every element touches in each analysis mode, and where the value comes
// from.  Follows crates/krets-solver/src/stampable.rs (push order preserved, because the reference sums duplicate
// triplets in push order — solver.rs:59-70 — and f64 addition is not associative).
#include <algorithm>
#include <cmath>
#include <complex>
#include <map>

#include "plan.hpp"

namespace krb {

namespace {
struct Emit {
    Stamps &s;
    void m(int r, int c, SrcKind k, bool neg, int e) {
        if (r >= 0 && c >= 0) s.mat.push_back({r, c, {k, neg, e}});
    }
    void b(int r, SrcKind k, bool neg, int e) {
        if (r >= 0) s.rhs.push_back({r, {k, neg, e}});
    }
    // group-1 two-terminal pattern (p,p)+y (m,m)+y (p,m)-y (m,p)-y   stampable.rs:229-244, 384-396, 456-465, 601-610
    void g1(int ip, int im, SrcKind k, int e) {
        m(ip, ip, k, false, e);
        m(im, im, k, false, e);
        if (ip >= 0 && im >= 0) {
            m(ip, im, k, true, e);
            m(im, ip, k, true, e);
        }
    }
};
}  // namespace

int build_stamps(const Circuit &c, int mode, int integrator, Stamps &out) {
    const bool trap = (mode == KRB_MODE_TRAN && integrator == KRB_INTEGRATOR_TRAP);
    out.mat.clear();
    out.rhs.clear();
    Emit E{out};
    for (int e = 0; e < (int)c.el.size(); e++) {
        const krb_element &el = c.el[e];
        const int ip = el.ip, im = el.im, ic = el.ic;
        switch (el.kind) {
        case KRB_R:   // stampable.rs:202-317 (g2 is never set by the parser, resistor.rs:47)
            E.g1(ip, im, S_G, e);
            break;
        case KRB_C:
            if (mode == KRB_MODE_DC) break;   // filtered out of OP/DC (op.rs:20-24, dc.rs:60-64)
            if (mode == KRB_MODE_AC) {        // :370-421
                if (!(el.flags & KRB_ELEM_FLAG_G2)) E.g1(ip, im, S_JWV, e);
                else {
                    if (ic >= 0) { E.m(ic, ip, S_JWV, true, e); E.m(ic, im, S_JWV, false, e); E.m(ic, ic, S_ONE, false, e); }
                }
            } else {                          // :442-505
                E.g1(ip, im, S_VOH, e);
                E.b(ip, S_CAP_HIST, false, e);   // b[ip] += -i_eq = +(C/h) v_prev
                if (trap) E.b(ip, S_CAP_IPREV, false, e);   // trapezoidal: ... + i_prev
                E.b(im, S_CAP_HIST, true, e);    // b[im] += +i_eq
                if (trap) E.b(im, S_CAP_IPREV, true, e);
            }
            break;
        case KRB_I:
            if (mode == KRB_MODE_AC) {        // :532-559 — no (ic,ic) entry: structurally empty row
                if (ic >= 0) { E.m(ip, ic, S_ONE, false, e); E.m(im, ic, S_ONE, true, e); }
            } else {                          // :509-530
                if (ic >= 0) { E.m(ic, ic, S_ONE, false, e); E.m(ip, ic, S_ONE, false, e); E.m(im, ic, S_ONE, true, e); }
            }
            E.b(ic, S_VAL, false, e);         // :561-584
            break;
        case KRB_D:
            E.g1(ip, im, S_DIODE_G, e);       // :588-641
            if (mode != KRB_MODE_AC) {        // :643-673
                E.b(ip, S_DIODE_IEQ, true, e);
                E.b(im, S_DIODE_IEQ, false, e);
            }
            break;
        case KRB_L:
            if (ic < 0) break;
            if (mode == KRB_MODE_DC) {        // :677-704
                E.m(ip, ic, S_ONE, false, e); E.m(im, ic, S_ONE, true, e);
                E.m(ic, ip, S_ONE, false, e); E.m(ic, im, S_ONE, true, e);
            } else if (mode == KRB_MODE_AC) { // :706-746
                E.m(ip, ic, S_ONE, false, e); E.m(im, ic, S_ONE, true, e);
                E.m(ic, ip, S_ONE, false, e); E.m(ic, im, S_ONE, true, e);
                E.m(ic, ic, S_JWV, true, e);
            } else {                          // :767-816
                E.m(ic, ic, S_VOH, true, e);
                E.m(ip, ic, S_ONE, false, e); E.m(ic, ip, S_ONE, false, e);
                E.m(im, ic, S_ONE, true, e);  E.m(ic, im, S_ONE, true, e);
                E.b(ic, S_IND_HIST, true, e);
                if (trap) E.b(ic, S_IND_VPREV, true, e);    // trapezoidal: b[ic] = -(2L/h) i_prev - v_prev
            }
            break;
        case KRB_NMOS:
            if (el.flags & KRB_ELEM_FLAG_LEVEL1) {
                // EXTENSION: the level-1 companion at the linearisation point, drain d = ip, gate g = im, source s = ic.
                // i_d enters the drain and leaves the source: rows d and s carry opposite signs (the reference's unreachable stamp,
                // stampable.rs:846-868, pushes (s,d) and (s,g) with the drain row's signs, which violates KCL at the source).
                // AC: the small-signal conductances at the operating point, no excitation (:897-912 returns none).
                const int d = ip, g = im, s = ic;
                E.m(d, d, S_MOS_GDS, false, e);
                E.m(s, s, S_MOS_GDSGM, false, e);
                if (d >= 0 && s >= 0) { E.m(d, s, S_MOS_GDSGM, true, e); E.m(s, d, S_MOS_GDS, true, e); }
                if (s >= 0 && g >= 0) E.m(s, g, S_MOS_GM, true, e);
                if (d >= 0 && g >= 0) E.m(d, g, S_MOS_GM, false, e);
                if (mode != KRB_MODE_AC) {   // :889-899 order: source row first
                    E.b(s, S_MOS_IEQ, false, e);
                    E.b(d, S_MOS_IEQ, true, e);
                }
                break;
            }
            if (mode == KRB_MODE_AC) return KRB_ERR_UNSUPPORTED;   // :914-921 todo!()
            break;                            // :842-844: index lookups never match -> stamps nothing
        case KRB_V:                           // :960-1050
            if (ic >= 0) {
                E.m(ip, ic, S_ONE, false, e); E.m(ic, ip, S_ONE, false, e);
                E.m(im, ic, S_ONE, true, e);  E.m(ic, im, S_ONE, true, e);
                E.b(ic, mode == KRB_MODE_DC ? S_VAL : mode == KRB_MODE_AC ? S_ACAMP : S_WAVE, false, e);
            }
            break;
        default:
            return KRB_ERR_UNSUPPORTED;
        }
    }
    return KRB_OK;
}

void build_pattern(int n, const Stamps &s, Pattern &p) {
    p.n = n;
    // unique keys in CSC order
    std::vector<std::pair<int64_t, int>> keyed(s.mat.size());
    for (size_t i = 0; i < s.mat.size(); i++) keyed[i] = {(int64_t)s.mat[i].col * n + s.mat[i].row, (int)i};
    std::stable_sort(keyed.begin(), keyed.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
    p.colptr.assign(n + 1, 0);
    p.rowidx.clear();
    p.contribs.clear();
    int64_t last = -1;
    for (auto &kv : keyed) {
        if (kv.first != last) {
            last = kv.first;
            int col = (int)(kv.first / n), row = (int)(kv.first % n);
            p.rowidx.push_back(row);
            p.contribs.emplace_back();
            p.colptr[col + 1]++;
        }
        p.contribs.back().push_back(kv.second);   // stable sort keeps stamp order inside an entry
    }
    for (int j = 0; j < n; j++) p.colptr[j + 1] += p.colptr[j];
    p.nnz = (int)p.rowidx.size();
    p.rhs_contribs.assign(n, {});
    for (size_t i = 0; i < s.rhs.size(); i++) p.rhs_contribs[s.rhs[i].row].push_back((int)i);
}

double transient_value_at(const krb_element &e, double time) {
    const double *w = e.wave;
    if (e.waveform == KRB_WAVE_PULSE) {   // voltage_source.rs:70-93
        const double v0 = w[0], v1 = w[1], td = w[2], tr = w[3], tf = w[4], pw = w[5], per = w[6];
        if (time < td) return v0;
        const double t = std::fmod(time - td, per);
        if (t < tr) return v0 + (v1 - v0) * (t / tr);
        if (t < tr + pw) return v1;
        if (t < tr + pw + tf) return v1 + (v0 - v1) * ((t - tr - pw) / tf);
        return v0;
    }
    if (e.waveform == KRB_WAVE_SIN) {     // voltage_source.rs:52-65
        const double off = w[0], amp = w[1], freq = w[2], delay = w[3], damp = w[4], phase = w[5];
        if (time < delay) return off;
        const double PI = 3.14159265358979323846264338327950288;
        const double omega = 2.0 * PI * freq;
        const double ph = phase * (PI / 180.0);
        const double t = time - delay;
        return off + amp * std::exp(-damp * t) * std::sin(omega * t + ph);
    }
    return e.value;                       // voltage_source.rs:250
}

static const double KB = 1.380649e-23, QE = 1.602176634e-19, TEMP = 300.0;   // constants.rs:1-3
static inline double thermal_voltage() { return KB * TEMP / QE; }            // constants.rs:4

void nominal_magnitudes(const Circuit &c, const Stamps &s, const Pattern &p, int mode, double h, double fnom, std::vector<double> &mag) {
    std::vector<std::complex<double>> val;
    nominal_values(c, s, p, mode, h, fnom, val);
    mag.assign(p.nnz, 0.0);
    for (int q = 0; q < p.nnz; q++) mag[q] = std::hypot(val[q].real(), val[q].imag());
}

// the signed (complex) nominal value of every entry: what the magnitudes above are taken of, and what the static pivot plan of
// the lane programs is checked against (compiler.cpp: static_pivots_ok)
void nominal_values(const Circuit &c, const Stamps &s, const Pattern &p, int mode, double h, double fnom, std::vector<std::complex<double>> &val) {
    const double PI = 3.14159265358979323846264338327950288;
    val.assign(p.nnz, std::complex<double>(0.0, 0.0));
    for (int q = 0; q < p.nnz; q++) {
        double re = 0.0, imv = 0.0;
        for (int ci : p.contribs[q]) {
            const Src &src = s.mat[ci].s;
            const krb_element &el = c.el[src.elem];
            double vr = 0.0, vi = 0.0;
            switch (src.kind) {
            case S_ONE: vr = 1.0; break;
            case S_G: vr = 1.0 / el.value; break;
            case S_VAL: vr = el.value; break;
            case S_VOH: vr = el.value / h; break;
            case S_JWV: vi = 2.0 * PI * fnom * el.value; break;
            case S_DIODE_G: {
                // the un-biased junction (vd = 0): g = Is / (n Vt) ~ 4e-11.  A diode conductance spans > 20 decades over the Newton
                // iterates and the lanes (reverse bias: -> 0), so the STATIC pivot plan must not rest on one wherever a resistor,
                // companion or source-coupling entry is available — the weighted matching then picks that entry, as faer's
                // per-factorisation partial pivoting (op.rs:50) would on every iterate where the diode is off.
                const double nvt = el.n * thermal_voltage();
                vr = el.is / nvt;
                break;
            }
            // level-1 NMOS one volt into saturation (v_ov = 1, v_ds = 1): like the diode, its conductances vanish in cut-off, so
            // the static pivot plan prefers any resistor / source-coupling entry of comparable size
            case S_MOS_GM: vr = el.wave[1] * (1.0 + el.wave[2]); break;
            case S_MOS_GDS: vr = 0.5 * el.wave[1] * el.wave[2]; break;
            case S_MOS_GDSGM: vr = el.wave[1] * (1.0 + el.wave[2]) + 0.5 * el.wave[1] * el.wave[2]; break;
            default: break;
            }
            if (src.neg) { vr = -vr; vi = -vi; }
            re += vr; imv += vi;
        }
        val[q] = std::complex<double>(re, imv);
    }
    (void)mode;
}

void evaluate_pool(const Plan &plan, const Circuit &c, double h, std::vector<double> &out) {
    out.resize(plan.pool.size());
    for (size_t i = 0; i < plan.pool.size(); i++) {
        const PoolEntry &pe = plan.pool[i];
        double v = 0.0;
        const krb_element *el = pe.elem >= 0 ? &c.el[pe.elem] : nullptr;
        switch (pe.kind) {
        case P_CONST: v = pe.c; break;
        case P_G: v = 1.0 / el->value; break;                       // stampable.rs:232
        case P_VAL: v = el->value; break;
        case P_VOH: v = el->value / h; break;                       // stampable.rs:449, 781
        case P_ACAMP: v = el->ac_amplitude; break;
        case P_DIO_ISNVT: v = el->is / (el->n * thermal_voltage()); break;   // diode.rs:47
        case P_DIO_NVT: v = el->n * thermal_voltage(); break;
        case P_DIO_IS: v = el->is; break;
        case P_MOS_VTH: v = el->wave[0]; break;
        case P_MOS_BETA: v = el->wave[1]; break;
        case P_MOS_LAMBDA: v = el->wave[2]; break;
        case P_DIO_VCRIT: {                                          // diode.rs:65-70
            // n * Vt * ln(f64::MAX * n * Vt / is), evaluated left to right as the reference does (f64::MAX * n may overflow)
            v = el->n * thermal_voltage() * std::log(1.7976931348623157e308 * el->n * thermal_voltage() / el->is);
            break;
        }
        }
        out[i] = pe.neg ? -v : v;
    }
}

}  // namespace krb
