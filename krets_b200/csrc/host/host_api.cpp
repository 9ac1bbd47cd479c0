// Host mirror of the krets public API over the C ABI (include/krets_b200_host.h).
#include <cmath>
#include <cstring>
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

struct krh_solver {                     // Solver { circuit, config }  solver.rs:17-20
    krh::ParsedCircuit pc;
    krb_config cfg;
    krb_circuit *kc = nullptr;
    std::map<int, std::unique_ptr<PlanHolder>> dc_plans;      // keyed by sweep element (-1: none)
    std::unique_ptr<PlanHolder> ac_plan;
    std::map<std::pair<double, int>, std::unique_ptr<PlanHolder>> tran_plans; // keyed by (time step, integrator)
    ~krh_solver() { dc_plans.clear(); ac_plan.reset(); tran_plans.clear(); krb_circuit_destroy(kc); }
};

struct krh_result {
    int kind = 0;
    int64_t rows = 0;
    int n = 0;
    std::vector<double> data, axis;
};

static int fail(int code, const std::string &msg) { g_herr = msg; return code; }
static int fail_krb(int code) { g_herr = krb_last_error_string(); return code; }

static krb_plan *get_dc_plan(krh_solver *s, int sweep_element, int &rc) {
    auto it = s->dc_plans.find(sweep_element);
    if (it != s->dc_plans.end()) { rc = KRB_OK; return it->second->p; }
    krb_plan_opts o;
    krb_default_plan_opts(&o);
    o.sweep_element = sweep_element;
    auto h = std::make_unique<PlanHolder>();
    rc = krb_analyze(s->kc, KRB_MODE_DC, &o, &h->p);
    if (rc) return nullptr;
    krb_plan *p = h->p;
    s->dc_plans[sweep_element] = std::move(h);
    return p;
}

// lane statuses -> the Err the reference would have returned at the first failing point
static int status_to_error(const std::vector<int32_t> &st) {
    for (int32_t v : st) {
        if (v & KRB_LANE_ZERO_PIVOT) return fail(KRB_ERR_DECOMPOSITION_FAILED, "Decomposition failed");
        if (v & KRB_LANE_NOT_CONVERGED) return fail(KRB_ERR_MAX_ITERATIONS, "Maximum iterations exceeded");
    }
    return KRB_OK;
}

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
    x.resize(n);
    return status_to_error(st);
}

int32_t krh_solver_solve_op(krh_solver *s, krh_result **out) {
    if (!s || !out) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    auto r = std::make_unique<krh_result>();
    int rc = solve_op_vec(s, r->data);
    if (rc) return rc;
    r->kind = KRH_RESULT_OP; r->rows = 1; r->n = (int)s->pc.unknown_names.size();
    *out = r.release();
    return KRB_OK;
}

// dc.rs:14-124
int32_t krh_solver_solve_dc(krh_solver *s, const char *element, double start, double stop, double step_size, krh_result **out) {
    if (!s || !out || !element) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    int ei = -1;
    for (size_t i = 0; i < s->pc.elements.size(); i++)
        if (s->pc.elements[i].identifier == element) { ei = (int)i; break; }          // dc.rs:21-24
    if (ei < 0) return fail(KRB_ERR_ELEMENT_NOT_FOUND, std::string("Element '") + element + "' not found in the netlist");
    const int kind = s->pc.pod[ei].kind;
    if (kind != KRB_V && kind != KRB_I) return fail(KRB_ERR_INVALID_ELEMENT_FORMAT, "DC sweep element must be a voltage or current source");   // dc.rs:30-38
    const int num = krh_dc_num_steps(start, stop, step_size);                          // dc.rs:44-45
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
    std::vector<int32_t> st(num, 0);
    rc = krb_solve_dc_sweep(p, &s->cfg, vals.data(), num, nullptr, r->data.data(), st.data(), nullptr);
    if (rc) return fail_krb(rc);
    if ((rc = status_to_error(st))) return rc;
    *out = r.release();
    return KRB_OK;
}

// ac.rs:17-109
int32_t krh_solver_solve_ac(krh_solver *s, int32_t sweep_kind, uint32_t points, double fstart, double fstop, krh_result **out) {
    if (!s || !out || sweep_kind < 0 || sweep_kind > 2) return fail(KRB_ERR_INVALID_ARGUMENT, "invalid argument");
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
        if (!s->ac_plan) {
            krb_plan_opts o;
            krb_default_plan_opts(&o);
            o.nominal_frequency = std::sqrt(freqs.front() * freqs.back());
            auto h = std::make_unique<PlanHolder>();
            rc = krb_analyze(s->kc, KRB_MODE_AC, &o, &h->p);
            if (rc) return fail_krb(rc);
            krb_plan_info pi;
            if (krb_plan_get_info(h->p, &pi) == KRB_OK && pi.executor == 2 && freqs.front() < o.nominal_frequency) {
                // front executor: static pivots prefer the diagonal where it is within 1e3 of the column maximum at the nominal
                // frequency; the inductor branch diagonal -j*w*L (stampable.rs:715) is smallest at the lowest frequency, so plan there
                krb_plan_destroy(h->p);
                h->p = nullptr;
                o.nominal_frequency = freqs.front();
                rc = krb_analyze(s->kc, KRB_MODE_AC, &o, &h->p);
                if (rc) return fail_krb(rc);
            }
            s->ac_plan = std::move(h);
        }
        std::vector<int32_t> st(freqs.size(), 0);
        rc = krb_solve_ac(s->ac_plan->p, &s->cfg, freqs.data(), (int64_t)freqs.size(), nullptr, op.data(), 1, r->data.data(), st.data(), nullptr);
        if (rc) return fail_krb(rc);
        if ((rc = status_to_error(st))) return rc;
    }
    *out = r.release();
    return KRB_OK;
}

// transient.rs:14-114
int32_t krh_solver_solve_transient(krh_solver *s, double time_step, double stop_time, krh_result **out) {
    return krh_solver_solve_transient_with(s, time_step, stop_time, KRB_INTEGRATOR_BE_KRETS, out);
}
int32_t krh_solver_solve_transient_with(krh_solver *s, double time_step, double stop_time, int32_t integrator, krh_result **out) {
    if (!s || !out) return fail(KRB_ERR_INVALID_ARGUMENT, "null argument");
    const int nsteps = krh_transient_num_steps(time_step, stop_time);                  // transient.rs:29
    int rc;
    krb_plan *pdc = get_dc_plan(s, -1, rc);
    if (!pdc) return fail_krb(rc);
    const auto key = std::make_pair(time_step, (int)integrator);
    auto it = s->tran_plans.find(key);
    if (it == s->tran_plans.end()) {
        krb_plan_opts o;
        krb_default_plan_opts(&o);
        o.time_step = time_step;
        o.integrator = integrator;
        auto h = std::make_unique<PlanHolder>();
        rc = krb_analyze(s->kc, KRB_MODE_TRAN, &o, &h->p);
        if (rc) return fail_krb(rc);
        it = s->tran_plans.emplace(key, std::move(h)).first;
    }
    const int n = (int)s->pc.unknown_names.size();
    auto r = std::make_unique<krh_result>();
    r->kind = KRH_RESULT_TRANSIENT; r->rows = nsteps + 1; r->n = n;
    r->data.assign((size_t)(nsteps + 1) * std::max(n, 1), 0.0);
    r->axis.resize(nsteps + 1);
    for (int k = 0; k <= nsteps; k++) r->axis[k] = (double)k * time_step;             // transient.rs:24,44
    std::vector<int32_t> st(1, 0);
    rc = krb_solve_transient(pdc, it->second->p, &s->cfg, nsteps, 1, nullptr, r->data.data(), st.data(), nullptr);
    if (rc) return fail_krb(rc);
    if ((rc = status_to_error(st))) return rc;
    *out = r.release();
    return KRB_OK;
}

int32_t krh_result_kind(const krh_result *r) { return r ? r->kind : -1; }
int64_t krh_result_rows(const krh_result *r) { return r ? r->rows : -1; }
int32_t krh_result_num_unknowns(const krh_result *r) { return r ? r->n : -1; }
const double *krh_result_data(const krh_result *r) { return r ? r->data.data() : nullptr; }
const double *krh_result_axis(const krh_result *r) { return (r && !r->axis.empty()) ? r->axis.data() : nullptr; }
void krh_result_free(krh_result *r) { delete r; }

}  // extern "C"
