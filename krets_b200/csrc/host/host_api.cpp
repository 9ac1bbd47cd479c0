// Host mirror of the krets public API over the C ABI (include/krets_b200_host.h).
#include <cmath>
#include <cstring>
#include <exception>
#include <new>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/krets_b200_host.h"
#include "netlist.hpp"

namespace {
thread_local std::string g_herr;
}

struct krh_circuit { krh::ParsedCircuit pc; };

struct PlanHolder {
    krb_plan *p = nullptr;
    ~PlanHolder() { krb_plan_destroy(p); }
};

struct PlanKey {                        // one cached plan per (mode, executor request, sweep element, step, integrator)
    int mode, exec, sweep_element, integrator;
    double time_step;
    bool operator<(const PlanKey &o) const {
        if (mode != o.mode) return mode < o.mode;
        if (exec != o.exec) return exec < o.exec;
        if (sweep_element != o.sweep_element) return sweep_element < o.sweep_element;
        if (integrator != o.integrator) return integrator < o.integrator;
        return time_step < o.time_step;
    }
};

struct krh_solver {                     // Solver { circuit, config }  solver.rs:17-20
    krh::ParsedCircuit pc;
    krb_config cfg;
    krb_circuit *kc = nullptr;
    std::map<PlanKey, std::unique_ptr<PlanHolder>> plans;
    ~krh_solver() { plans.clear(); krb_circuit_destroy(kc); }
};

struct krh_result {
    int kind = 0;
    int64_t rows = 0;
    int n = 0;
    std::vector<double> data, axis;
};

static int fail(int code, const std::string &msg) { g_herr = msg; return code; }
static int fail_krb(int code) { g_herr = krb_last_error_string(); return code; }

// cached krb_analyze; `nominal_frequency` is only used when the plan is created (AC plans are keyed without it: one per solver)
static krb_plan *get_plan(krh_solver *s, int mode, int exec, int sweep_element, double time_step, int integrator, double nominal_frequency, int &rc,
                          int refine_steps = 0) {
    const PlanKey key{mode, exec + 16 * refine_steps, sweep_element, integrator, time_step};
    auto it = s->plans.find(key);
    if (it != s->plans.end()) { rc = KRB_OK; return it->second->p; }
    krb_plan_opts o;
    krb_default_plan_opts(&o);
    o.exec = exec;
    o.sweep_element = sweep_element;
    o.time_step = time_step;
    o.integrator = integrator;
    o.nominal_frequency = nominal_frequency;
    o.refine_steps = refine_steps;
    auto h = std::make_unique<PlanHolder>();
    rc = krb_analyze(s->kc, mode, &o, &h->p);
    if (rc) return nullptr;
    krb_plan *p = h->p;
    s->plans[key] = std::move(h);
    return p;
}
static krb_plan *get_dc_plan(krh_solver *s, int sweep_element, int &rc, int exec = KRB_EXEC_FUSED, int refine_steps = 0) {
    return get_plan(s, KRB_MODE_DC, exec, sweep_element, 0.0, 0, 0.0, rc, refine_steps);
}
// pivoting fallback: front executor (threshold pivoting inside the fronts, per lane) + 2 refinement steps + backward-error check
static const int KRH_FALLBACK_REFINE = 2;
static int plan_executor(const krb_plan *p) {
    krb_plan_info pi;
    return krb_plan_get_info(p, &pi) == KRB_OK ? pi.executor : -1;
}

// lane statuses -> the Err the reference would have returned at the first failing point.  The reference pivots per factorisation
// (faer, op.rs:50): a lane that still carries KRB_LANE_INACCURATE after the pivoting fallback, or that delivered inf/nan, is
// reported as a failed decomposition rather than returned as a success with wrong digits.
static int status_to_error(const std::vector<int32_t> &st) {
    for (int32_t v : st) {
        if (v & KRB_LANE_ZERO_PIVOT) return fail(KRB_ERR_DECOMPOSITION_FAILED, "Decomposition failed");
        if (v & KRB_LANE_NOT_CONVERGED) return fail(KRB_ERR_MAX_ITERATIONS, "Maximum iterations exceeded");
        if (v & KRB_LANE_NONFINITE) return fail(KRB_ERR_DECOMPOSITION_FAILED, "Decomposition failed (non-finite solution)");
        if (v & KRB_LANE_INACCURATE) return fail(KRB_ERR_DECOMPOSITION_FAILED, "Decomposition failed (static pivot growth >= 2^24 and no pivoting fallback for this circuit)");
    }
    return KRB_OK;
}
static std::vector<int64_t> inaccurate_lanes(const std::vector<int32_t> &st) {
    std::vector<int64_t> v;
    for (size_t i = 0; i < st.size(); i++) if ((st[i] & KRB_LANE_INACCURATE) && !(st[i] & ~(KRB_LANE_INACCURATE))) v.push_back((int64_t)i);
    return v;
}
static bool circuit_is_linear(const krh_solver *s) {
    for (const krb_element &e : s->pc.pod) if (e.kind == KRB_D || e.kind == KRB_NMOS) return false;
    return true;
}
// entry points must not let a C++ exception cross the C boundary (std::bad_alloc on absurd step counts)
#define KRH_GUARD_BEGIN try {
#define KRH_GUARD_END                                                                                    \
    } catch (const std::bad_alloc &) { return fail(KRB_ERR_OUT_OF_MEMORY, "out of host memory"); }       \
    catch (const std::exception &e) { return fail(KRB_ERR_INTERNAL, e.what()); }
static const int64_t KRH_MAX_ROWS = 0x3fffffff;   // krb_solve_transient's own bound on n_steps

extern "C" {

const char *krh_last_error_string(void) { return g_herr.c_str(); }

int32_t krh_parse_circuit_description(const char *text, krh_circuit **out) {
    if (!text || !out) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    auto c = std::make_unique<krh_circuit>();
    std::string err;
    int rc = krh::parse_circuit_description(text, c->pc, err);
    if (rc) return fail(rc, err);
    *out = c.release();
    return KRB_OK;
}
int32_t krh_parse_circuit_description_file(const char *path, krh_circuit **out) {   // parser.rs:156-164
    if (!path || !out) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(KRH_ERR_IO, std::string("cannot open ") + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return krh_parse_circuit_description(ss.str().c_str(), out);
}
void krh_circuit_free(krh_circuit *c) { delete c; }
int32_t krh_circuit_num_unknowns(const krh_circuit *c) { return c ? (int32_t)c->pc.unknown_names.size() : -1; }
const char *krh_circuit_unknown_name(const krh_circuit *c, int32_t i) {
    return (c && i >= 0 && i < (int32_t)c->pc.unknown_names.size()) ? c->pc.unknown_names[i].c_str() : nullptr;
}
int32_t krh_circuit_index_of(const krh_circuit *c, const char *key) {
    if (!c || !key) return -1;
    auto it = c->pc.index_map.find(key);
    return it == c->pc.index_map.end() ? -1 : it->second;
}
int32_t krh_circuit_num_elements(const krh_circuit *c) { return c ? (int32_t)c->pc.elements.size() : -1; }
const char *krh_circuit_element_identifier(const krh_circuit *c, int32_t e) {
    return (c && e >= 0 && e < (int32_t)c->pc.elements.size()) ? c->pc.elements[e].identifier.c_str() : nullptr;
}
const krb_element *krh_circuit_elements(const krh_circuit *c) { return c ? c->pc.pod.data() : nullptr; }
const uint8_t *krh_circuit_is_current(const krh_circuit *c) { return c ? c->pc.is_current.data() : nullptr; }
int32_t krh_circuit_set_device_models(krh_circuit *c, int32_t models) {
    if (!c || (models != KRH_DEVICES_REFERENCE && models != KRH_DEVICES_LEVEL1)) return -1;
    int32_t k = 0;
    for (krb_element &e : c->pc.pod)
        if (e.kind == KRB_NMOS) {
            if (models == KRH_DEVICES_LEVEL1) e.flags |= KRB_ELEM_FLAG_LEVEL1; else e.flags &= ~KRB_ELEM_FLAG_LEVEL1;
            k++;
        }
    return k;
}

int64_t krh_generate_frequencies(int32_t kind, uint32_t points, double fstart, double fstop, double *out, int64_t cap) {
    if (kind < 0 || kind > 2) return -1;
    std::vector<double> f = krh::generate_frequencies(kind, points, fstart, fstop);
    if (out) {
        if ((int64_t)f.size() > cap) return -(int64_t)f.size();
        if (!f.empty()) std::memcpy(out, f.data(), f.size() * 8);
    }
    return (int64_t)f.size();
}
int32_t krh_dc_num_steps(double start, double stop, double step_size) {   // `as usize` truncates/saturates
    const double v = std::fabs((stop - start) / step_size);
    if (!(v == v)) return 1;
    if (v > 2147483000.0) return 2147483647;
    return (int32_t)v + 1;
}
int32_t krh_transient_num_steps(double time_step, double stop_time) {
    const double v = std::round(stop_time / time_step);
    if (!(v == v) || v < 0) return 0;
    if (v > 2147483000.0) return 2147483647;
    return (int32_t)v;
}

int32_t krh_solver_new(const krh_circuit *c, const krb_config *cfg, krh_solver **out) {
    if (!c || !out) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    auto s = std::make_unique<krh_solver>();
    s->pc = c->pc;   // Solver owns its Circuit by value
    if (cfg) s->cfg = *cfg; else krb_default_config(&s->cfg);
    int rc = krb_circuit_create(s->pc.pod.data(), (int32_t)s->pc.pod.size(), (int32_t)s->pc.unknown_names.size(), s->pc.is_current.data(), &s->kc);
    if (rc) return fail_krb(rc);
    *out = s.release();
    return KRB_OK;
}
void krh_solver_free(krh_solver *s) { delete s; }

// op.rs:15-86
static int solve_op_vec(krh_solver *s, std::vector<double> &x) {
    int rc;
    krb_plan *p = get_dc_plan(s, -1, rc);
    if (!p) return fail_krb(rc);
    const int n = (int)s->pc.unknown_names.size();
    x.assign(std::max(n, 1), 0.0);
    std::vector<int32_t> st(1, 0);
    rc = krb_solve_op(p, &s->cfg, 1, nullptr, x.data(), nullptr, st.data(), nullptr);
    if (rc) return fail_krb(rc);
    if (!inaccurate_lanes(st).empty() && circuit_is_linear(s)) {
        // pivoting fallback: the front executor pivots inside its fronts per lane (front_kernels.cuh)
        krb_plan *pf = get_dc_plan(s, -1, rc, KRB_EXEC_STAGED, KRH_FALLBACK_REFINE);
        if (!pf) return fail_krb(rc);
        st[0] = 0;
        rc = krb_solve_op(pf, &s->cfg, 1, nullptr, x.data(), nullptr, st.data(), nullptr);
        if (rc) return fail_krb(rc);
    }
    x.resize(n);
    return status_to_error(st);
}

int32_t krh_solver_solve_op(krh_solver *s, krh_result **out) {
    if (!s || !out) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    KRH_GUARD_BEGIN
    auto r = std::make_unique<krh_result>();
    int rc = solve_op_vec(s, r->data);
    if (rc) return rc;
    r->kind = KRH_RESULT_OP; r->rows = 1; r->n = (int)s->pc.unknown_names.size();
    *out = r.release();
    return KRB_OK;
    KRH_GUARD_END
}

// dc.rs:14-124
int32_t krh_solver_solve_dc(krh_solver *s, const char *element, double start, double stop, double step_size, krh_result **out) {
    return krh_solver_solve_dc_with(s, element, start, stop, step_size, KRH_DC_REFERENCE, out);
}
int32_t krh_solver_solve_dc_with(krh_solver *s, const char *element, double start, double stop, double step_size, int32_t mode, krh_result **out) {
    if (!s || !out || !element) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    if (mode != KRH_DC_REFERENCE && mode != KRH_DC_PARALLEL_COLD_START) return fail(KRB_ERR_INVALID_ARGUMENT, "DC sweep: unknown mode");
    int ei = -1;
    for (size_t i = 0; i < s->pc.elements.size(); i++)
        if (s->pc.elements[i].identifier == element) { ei = (int)i; break; }          // dc.rs:21-24
    if (ei < 0) return fail(KRB_ERR_ELEMENT_NOT_FOUND, std::string("Element '") + element + "' not found in the netlist");
    const int kind = s->pc.pod[ei].kind;
    if (kind != KRB_V && kind != KRB_I) return fail(KRB_ERR_INVALID_ELEMENT_FORMAT, "DC sweep element must be a voltage or current source");   // dc.rs:30-38
    if (!std::isfinite(start) || !std::isfinite(stop) || !std::isfinite(step_size) || step_size == 0.0)
        return fail(KRB_ERR_INVALID_ARGUMENT, "DC sweep: start, stop and a non-zero step_size must be finite");
    const int num = krh_dc_num_steps(start, stop, step_size);                          // dc.rs:44-45
    if ((int64_t)num > KRH_MAX_ROWS) return fail(KRB_ERR_INVALID_ARGUMENT, "DC sweep: too many sweep points");
    KRH_GUARD_BEGIN
    int rc;
    krb_plan *p = get_dc_plan(s, ei, rc);
    if (!p) return fail_krb(rc);
    const int n = (int)s->pc.unknown_names.size();
    auto r = std::make_unique<krh_result>();
    r->kind = KRH_RESULT_DC; r->rows = num; r->n = n;
    r->data.assign((size_t)num * std::max(n, 1), 0.0);
    r->axis.resize(num);
    std::vector<double> vals(num);
    for (int i = 0; i < num; i++) { vals[i] = start + ((double)i * step_size); r->axis[i] = (double)i; }   // dc.rs:48, :97
    if (!circuit_is_linear(s) && mode == KRH_DC_REFERENCE) {
        // dc.rs:58,112: with Newton-Raphson every sweep point starts from its predecessor's solution, so the points are sequential:
        // one lane, the sweep points as its rows (x[0][row][unknown] is exactly the result layout)
        int32_t st1 = 0;
        rc = krb_solve_dc_sweep_sequential(p, &s->cfg, start, step_size, num, 1, nullptr, r->data.data(), nullptr, &st1, nullptr);
        if (rc) return fail_krb(rc);
        std::vector<int32_t> sv(1, st1);
        if ((rc = status_to_error(sv))) return rc;
        *out = r.release();
        return KRB_OK;
    }
    std::vector<int32_t> st(num, 0);
    rc = krb_solve_dc_sweep(p, &s->cfg, vals.data(), num, nullptr, r->data.data(), st.data(), nullptr);
    if (rc) return fail_krb(rc);
    const std::vector<int64_t> redo = inaccurate_lanes(st);
    if (!redo.empty() && circuit_is_linear(s)) {            // pivoting fallback for the flagged points only
        krb_plan *pf = get_dc_plan(s, ei, rc, KRB_EXEC_STAGED, KRH_FALLBACK_REFINE);
        if (!pf) return fail_krb(rc);
        std::vector<double> v2(redo.size()), x2(redo.size() * (size_t)std::max(n, 1));
        std::vector<int32_t> st2(redo.size(), 0);
        for (size_t i = 0; i < redo.size(); i++) v2[i] = vals[redo[i]];
        rc = krb_solve_dc_sweep(pf, &s->cfg, v2.data(), (int64_t)redo.size(), nullptr, x2.data(), st2.data(), nullptr);
        if (rc) return fail_krb(rc);
        for (size_t i = 0; i < redo.size(); i++) {
            std::memcpy(&r->data[(size_t)redo[i] * n], &x2[i * (size_t)n], (size_t)n * 8);
            st[redo[i]] = st2[i];
        }
    }
    if ((rc = status_to_error(st))) return rc;
    *out = r.release();
    return KRB_OK;
    KRH_GUARD_END
}

// ac.rs:17-109
int32_t krh_solver_solve_ac(krh_solver *s, int32_t sweep_kind, uint32_t points, double fstart, double fstop, krh_result **out) {
    if (!s || !out || sweep_kind < 0 || sweep_kind > 2) return fail(KRB_ERR_INVALID_ARGUMENT, "invalid argument");
    KRH_GUARD_BEGIN
    std::vector<double> op;
    int rc = solve_op_vec(s, op);                                                     // ac.rs:25
    if (rc) return rc;
    std::vector<double> all = krh::generate_frequencies(sweep_kind, points, fstart, fstop), freqs;
    for (double f : all) if (!(f <= 0.0)) freqs.push_back(f);                          // ac.rs:40-45
    const int n = (int)s->pc.unknown_names.size();
    auto r = std::make_unique<krh_result>();
    r->kind = KRH_RESULT_AC; r->rows = (int64_t)freqs.size(); r->n = n;
    r->axis = freqs;
    r->data.assign(freqs.size() * (size_t)std::max(n, 1) * 2, 0.0);
    if (!freqs.empty()) {
        double fnom = std::sqrt(freqs.front() * freqs.back());
        krb_plan *pa = get_plan(s, KRB_MODE_AC, KRB_EXEC_FUSED, -1, 0.0, 0, fnom, rc);
        if (!pa) return fail_krb(rc);
        if (plan_executor(pa) == 2 && freqs.front() < fnom) {
            // front executor: static pivots prefer the diagonal where it is within 1e3 of the column maximum at the nominal
            // frequency; the inductor branch diagonal -j*w*L (stampable.rs:715) is smallest at the lowest frequency, so plan there
            pa = get_plan(s, KRB_MODE_AC, KRB_EXEC_STAGED, -1, 0.0, 0, freqs.front(), rc);
            if (!pa) return fail_krb(rc);
        }
        std::vector<int32_t> st(freqs.size(), 0);
        rc = krb_solve_ac(pa, &s->cfg, freqs.data(), (int64_t)freqs.size(), nullptr, op.data(), 1, r->data.data(), st.data(), nullptr);
        if (rc) return fail_krb(rc);
        const std::vector<int64_t> redo = inaccurate_lanes(st);
        if (!redo.empty() && circuit_is_linear(s)) {        // pivoting fallback for the flagged frequencies only
            std::vector<double> f2(redo.size()), x2(redo.size() * (size_t)std::max(n, 1) * 2);
            std::vector<int32_t> st2(redo.size(), 0);
            for (size_t i = 0; i < redo.size(); i++) f2[i] = freqs[redo[i]];
            krb_plan *pf = get_plan(s, KRB_MODE_AC, KRB_EXEC_STAGED, -1, 0.0, 0, f2.front(), rc, KRH_FALLBACK_REFINE);
            if (!pf) return fail_krb(rc);
            rc = krb_solve_ac(pf, &s->cfg, f2.data(), (int64_t)redo.size(), nullptr, op.data(), 1, x2.data(), st2.data(), nullptr);
            if (rc) return fail_krb(rc);
            for (size_t i = 0; i < redo.size(); i++) {
                std::memcpy(&r->data[(size_t)redo[i] * n * 2], &x2[i * (size_t)n * 2], (size_t)n * 16);
                st[redo[i]] = st2[i];
            }
        }
        if ((rc = status_to_error(st))) return rc;
    }
    *out = r.release();
    return KRB_OK;
    KRH_GUARD_END
}

// transient.rs:14-114
int32_t krh_solver_solve_transient(krh_solver *s, double time_step, double stop_time, krh_result **out) {
    return krh_solver_solve_transient_with(s, time_step, stop_time, KRB_INTEGRATOR_BE_KRETS, out);
}
int32_t krh_solver_solve_transient_with(krh_solver *s, double time_step, double stop_time, int32_t integrator, krh_result **out) {
    if (!s || !out) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    if (!std::isfinite(time_step) || !std::isfinite(stop_time) || !(time_step > 0.0))
        return fail(KRB_ERR_INVALID_ARGUMENT, "transient: time_step must be positive and finite, stop_time finite");
    const int nsteps = krh_transient_num_steps(time_step, stop_time);                  // transient.rs:29
    if ((int64_t)nsteps + 1 > KRH_MAX_ROWS) return fail(KRB_ERR_INVALID_ARGUMENT, "transient: too many time steps");
    KRH_GUARD_BEGIN
    int rc;
    // The DC and the step plan choose their executor independently (their patterns differ: capacitors are dropped and inductor
    // branches have no diagonal in DC); the runtime needs both on the same side, so when exactly one of them selected the front
    // executor the other one is analysed for it as well.
    krb_plan *pdc = get_dc_plan(s, -1, rc);
    if (!pdc) return fail_krb(rc);
    krb_plan *ptr = get_plan(s, KRB_MODE_TRAN, KRB_EXEC_FUSED, -1, time_step, integrator, 0.0, rc);
    if (!ptr) return fail_krb(rc);
    const bool fdc = plan_executor(pdc) == 2, ftr = plan_executor(ptr) == 2;
    if (fdc != ftr) {
        if (!fdc) pdc = get_dc_plan(s, -1, rc, KRB_EXEC_STAGED);
        else ptr = get_plan(s, KRB_MODE_TRAN, KRB_EXEC_STAGED, -1, time_step, integrator, 0.0, rc);
        if (!pdc || !ptr) return fail_krb(rc);
    }
    const int n = (int)s->pc.unknown_names.size();
    auto r = std::make_unique<krh_result>();
    r->kind = KRH_RESULT_TRANSIENT; r->rows = nsteps + 1; r->n = n;
    r->data.assign((size_t)(nsteps + 1) * std::max(n, 1), 0.0);
    r->axis.resize(nsteps + 1);
    for (int k = 0; k <= nsteps; k++) r->axis[k] = (double)k * time_step;             // transient.rs:24,44
    std::vector<int32_t> st(1, 0);
    rc = krb_solve_transient(pdc, ptr, &s->cfg, nsteps, 1, nullptr, r->data.data(), st.data(), nullptr);
    if (rc) return fail_krb(rc);
    if (!inaccurate_lanes(st).empty() && circuit_is_linear(s)) {       // pivoting fallback: both plans on the front executor
        pdc = get_dc_plan(s, -1, rc, KRB_EXEC_STAGED, KRH_FALLBACK_REFINE);
        if (!pdc) return fail_krb(rc);
        ptr = get_plan(s, KRB_MODE_TRAN, KRB_EXEC_STAGED, -1, time_step, integrator, 0.0, rc, KRH_FALLBACK_REFINE);
        if (!ptr) return fail_krb(rc);
        st[0] = 0;
        rc = krb_solve_transient(pdc, ptr, &s->cfg, nsteps, 1, nullptr, r->data.data(), st.data(), nullptr);
        if (rc) return fail_krb(rc);
    }
    if ((rc = status_to_error(st))) return rc;
    *out = r.release();
    return KRB_OK;
    KRH_GUARD_END
}

int32_t krh_result_kind(const krh_result *r) { return r ? r->kind : -1; }
int64_t krh_result_rows(const krh_result *r) { return r ? r->rows : -1; }
int32_t krh_result_num_unknowns(const krh_result *r) { return r ? r->n : -1; }
const double *krh_result_data(const krh_result *r) { return r ? r->data.data() : nullptr; }
const double *krh_result_axis(const krh_result *r) { return (r && !r->axis.empty()) ? r->axis.data() : nullptr; }
void krh_result_free(krh_result *r) { delete r; }

}  // extern "C"
