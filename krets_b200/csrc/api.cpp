// C ABI entry points (include/krets_b200.h).  Host-only parts (circuit table, symbolic analysis, introspection) work
// without a GPU; everything numeric goes to the CUDA runtime and fails loudly without a device.
#include <cstring>
#include <new>
#include <string>

#include "plan.hpp"
#include "runtime.hpp"
#include "vm_jit.hpp"

using namespace krb;

struct krb_circuit { Circuit c; };
struct krb_plan { Plan p; };

extern "C" {

void krb_default_config(krb_config *cfg) {   // config.rs:26-36
    cfg->relative_tolerance = 0.001;
    cfg->current_absolute_tolerance = 1e-12;
    cfg->voltage_absolute_tolerance = 1e-6;
    cfg->maximum_iterations = 300;
}
void krb_default_plan_opts(krb_plan_opts *o) {
    std::memset(o, 0, sizeof(*o));
    o->struct_size = (int32_t)sizeof(*o);
    o->exec = KRB_EXEC_FUSED;
    o->integrator = KRB_INTEGRATOR_BE_KRETS;
    o->sweep_element = -1;
}
void krb_default_io(krb_io *io) {
    std::memset(io, 0, sizeof(*io));
    io->struct_size = (int32_t)sizeof(*io);
    io->memory = KRB_MEM_HOST;
    io->layout = KRB_LAYOUT_POINT_MAJOR;
}

int32_t krb_circuit_create(const krb_element *elements, int32_t n_elements, int32_t n_unknowns, const uint8_t *unknown_is_current,
                           krb_circuit **out) {
    if (!out || n_elements < 0 || n_unknowns < 0 || (n_elements > 0 && !elements) || (n_unknowns > 0 && !unknown_is_current)) {
        set_error("krb_circuit_create: invalid argument");
        return KRB_ERR_INVALID_ARGUMENT;
    }
    krb_circuit *c = new (std::nothrow) krb_circuit();
    if (!c) return KRB_ERR_OUT_OF_MEMORY;
    c->c.el.assign(elements, elements + n_elements);
    c->c.n = n_unknowns;
    c->c.is_current.assign(unknown_is_current, unknown_is_current + n_unknowns);
    c->c.nonlinear = false;
    for (const auto &e : c->c.el) {
        if (e.kind < KRB_R || e.kind > KRB_NMOS) { delete c; set_error("unknown element kind"); return KRB_ERR_INVALID_ARGUMENT; }
        if (e.ip >= n_unknowns || e.im >= n_unknowns || e.ic >= n_unknowns || e.ip < -1 || e.im < -1 || e.ic < -1) {
            delete c; set_error("element index out of range"); return KRB_ERR_INVALID_ARGUMENT;
        }
        if (e.kind == KRB_D || e.kind == KRB_NMOS) c->c.nonlinear = true;   // elements.rs:143-148
    }
    *out = c;
    return KRB_OK;
}
void krb_circuit_destroy(krb_circuit *c) { delete c; }

int32_t krb_analyze(const krb_circuit *c, int32_t mode, const krb_plan_opts *opts, krb_plan **out) {
    if (!c || !out || mode < KRB_MODE_DC || mode > KRB_MODE_TRAN) { set_error("krb_analyze: invalid argument"); return KRB_ERR_INVALID_ARGUMENT; }
    krb_plan_opts o;
    krb_default_plan_opts(&o);
    if (opts) {
        if (opts->struct_size != (int32_t)sizeof(krb_plan_opts)) { set_error("krb_plan_opts.struct_size mismatch"); return KRB_ERR_INVALID_ARGUMENT; }
        o = *opts;
    }
    if (o.n_varied < 0 || (o.n_varied > 0 && !o.varied_elements)) { set_error("krb_analyze: varied_elements missing"); return KRB_ERR_INVALID_ARGUMENT; }
    if (o.exec != KRB_EXEC_FUSED && o.exec != KRB_EXEC_LANE_VM && o.exec != KRB_EXEC_STAGED) { set_error("unknown krb_plan_opts.exec"); return KRB_ERR_INVALID_ARGUMENT; }
    krb_plan *p = new (std::nothrow) krb_plan();
    if (!p) return KRB_ERR_OUT_OF_MEMORY;
    std::string err;
    int rc;
    NvtxRange nvtx_analyze(mode == KRB_MODE_DC ? "krb_analyze (DC): stamps, pivot plan, symbolic LU, executor tables"
                           : mode == KRB_MODE_AC ? "krb_analyze (AC): stamps, pivot plan, symbolic LU, executor tables"
                                                 : "krb_analyze (TRAN): stamps, pivot plan, symbolic LU, executor tables");
    try {
        rc = analyze(c->c, mode, o, p->p, err);
    } catch (const std::exception &ex) {
        rc = KRB_ERR_INTERNAL;
        err = ex.what();
    }
    if (rc) { set_error(err); delete p; return rc; }
    *out = p;
    return KRB_OK;
}
void krb_plan_destroy(krb_plan *p) {
    if (!p) return;
    device_state_free(p->p.dev);
    front_device_free(p->p.fdev);
    delete p;
}

int32_t krb_plan_get_info(const krb_plan *p, krb_plan_info *info) {
    if (!p || !info) return KRB_ERR_INVALID_ARGUMENT;
    const Plan &P = p->p;
    info->n = P.circuit.n;
    info->nnz = P.pattern.nnz;
    info->nnz_lu = (int32_t)P.sym.nnz_lu;
    info->mode = P.mode;
    info->nonlinear = P.circuit.nonlinear ? 1 : 0;
    info->scratch_slots = P.n_slots;
    info->spill_planes = P.n_planes;
    info->n_inputs = P.n_inputs;
    info->program_ops = P.sec[0].n_ops + P.sec[1].n_ops + P.sec[2].n_ops;
    info->flops_per_point = P.flops_per_point;
    info->hbm_bytes_per_point = P.hbm_bytes_per_point;
    info->alg_bytes_per_point = P.alg_bytes_per_point;
    info->executor = P.use_front ? 2 : P.use_chain ? 1 : 0;
    info->reserved = 0;
    return KRB_OK;
}
int32_t krb_plan_get_pattern(const krb_plan *p, int32_t *colptr, int32_t *rowidx) {
    if (!p || !colptr || !rowidx) return KRB_ERR_INVALID_ARGUMENT;
    const Pattern &pt = p->p.pattern;
    for (int j = 0; j <= pt.n; j++) colptr[j] = pt.colptr[j];
    for (int q = 0; q < pt.nnz; q++) rowidx[q] = pt.rowidx[q];
    return KRB_OK;
}
int32_t krb_plan_get_pivot_order(const krb_plan *p, int32_t *row_order, int32_t *col_order) {
    if (!p || !row_order || !col_order) return KRB_ERR_INVALID_ARGUMENT;
    const Symbolic &s = p->p.sym;
    for (int k = 0; k < s.n; k++) { row_order[k] = s.row_order[k]; col_order[k] = s.col_order[k]; }
    return KRB_OK;
}
int64_t krb_plan_jit_source(const krb_plan *p, int32_t host_mode, char *buf, int64_t cap) {
    if (!p || cap < 0 || (cap > 0 && !buf)) return -1;
    if (!vm_jit_eligible(p->p)) return 0;
    const std::string src = vm_jit_source(p->p, vm_params_size(), host_mode != 0);
    if (buf && cap > 0) {
        const size_t k = std::min<size_t>(src.size(), (size_t)cap - 1);
        std::memcpy(buf, src.data(), k);
        buf[k] = '\0';
    }
    return (int64_t)src.size() + 1;
}
int32_t krb_plan_jit_compile(const krb_plan *p, int64_t *cubin_bytes) {
    if (!p) return KRB_ERR_INVALID_ARGUMENT;
    if (cubin_bytes) *cubin_bytes = 0;
    if (!vm_jit_eligible(p->p)) { set_error("the plan's lane program is not eligible for a specialised kernel (executor or length)"); return KRB_ERR_UNSUPPORTED; }
    std::string cubin, err;
    if (!vm_jit_compile(vm_jit_source(p->p, vm_params_size(), false), cubin, err)) { set_error(err); return KRB_ERR_INTERNAL; }
    if (cubin_bytes) *cubin_bytes = (int64_t)cubin.size();
    return KRB_OK;
}
int64_t krb_plan_program_size(const krb_plan *p, int32_t section) {
    if (!p || section < 0 || section > 2) return -1;
    return (int64_t)p->p.sec[section].ops.size() / 2;
}
int32_t krb_plan_get_program(const krb_plan *p, int32_t section, uint32_t *ops) {
    if (!p || section < 0 || section > 2 || !ops) return KRB_ERR_INVALID_ARGUMENT;
    const auto &v = p->p.sec[section].ops;
    if (!v.empty()) std::memcpy(ops, v.data(), v.size() * sizeof(uint32_t));
    return KRB_OK;
}
/* introspection used by the CPU-side program emulator in tests/ (not part of the drop-in surface) */
int64_t krb_plan_pool_size(const krb_plan *p) { return p ? (int64_t)p->p.pool.size() : -1; }
int32_t krb_plan_eval_pool(const krb_plan *p, double *out) {
    if (!p || !out) return KRB_ERR_INVALID_ARGUMENT;
    std::vector<double> v;
    evaluate_pool(p->p, p->p.circuit, p->p.h_comp, v);
    if (!v.empty()) std::memcpy(out, v.data(), v.size() * 8);
    return KRB_OK;
}
int32_t krb_plan_wave_count(const krb_plan *p) { return p ? (int32_t)p->p.wave_elems.size() : -1; }
int32_t krb_plan_wave_value(const krb_plan *p, int32_t k, double time, double *out) {
    if (!p || k < 0 || k >= (int32_t)p->p.wave_elems.size() || !out) return KRB_ERR_INVALID_ARGUMENT;
    *out = transient_value_at(p->p.circuit.el[p->p.wave_elems[k]], time);
    return KRB_OK;
}
int32_t krb_plan_input_layout(const krb_plan *p, int32_t *input_sweep, int32_t *input_freq, int32_t *input_of_varied) {
    if (!p) return KRB_ERR_INVALID_ARGUMENT;
    if (input_sweep) *input_sweep = p->p.input_sweep;
    if (input_freq) *input_freq = p->p.input_freq;
    if (input_of_varied) for (size_t i = 0; i < p->p.input_of_varied.size(); i++) input_of_varied[i] = p->p.input_of_varied[i];
    return KRB_OK;
}
int64_t krb_plan_chain_size(const krb_plan *p) { return (p && p->p.chain.valid) ? (int64_t)p->p.chain.stream.size() : 0; }   /* 8-byte words */
int32_t krb_plan_get_chain(const krb_plan *p, uint32_t *rowinfo /*n+1*/, uint64_t *stream, int32_t *unknown_of_row) {
    if (!p || !p->p.chain.valid) return KRB_ERR_INVALID_ARGUMENT;
    const ChainPlan &ch = p->p.chain;
    if (rowinfo) std::memcpy(rowinfo, ch.rowinfo.data(), ch.rowinfo.size() * 4);
    if (stream) std::memcpy(stream, ch.stream.data(), ch.stream.size() * 8);
    if (unknown_of_row) std::memcpy(unknown_of_row, ch.unknown_of_row.data(), ch.unknown_of_row.size() * sizeof(int));
    return KRB_OK;
}
/* block-partitioned chain kernel tables (tests emulate the kernel on CPU): info = {valid, T, M, n_dict}; rows 2*T*M words; dict 4|8 words each */
int32_t krb_plan_chain_block_info(const krb_plan *p, int32_t *info /*4*/, double *nominal_diff) {
    if (!p || !p->p.chain.valid || !info) return KRB_ERR_INVALID_ARGUMENT;
    const ChainPlan::Block &b = p->p.chain.block;
    info[0] = b.valid ? 1 : 0; info[1] = b.T; info[2] = b.M; info[3] = b.n_dict;
    if (nominal_diff) *nominal_diff = b.nominal_diff;
    if (!b.valid) set_error(b.why_not);
    return KRB_OK;
}
int32_t krb_plan_get_chain_block(const krb_plan *p, uint32_t *rows, uint64_t *dict) {
    if (!p || !p->p.chain.valid || !p->p.chain.block.valid) return KRB_ERR_INVALID_ARGUMENT;
    const ChainPlan::Block &b = p->p.chain.block;
    if (rows) std::memcpy(rows, b.rows.data(), b.rows.size() * 4);
    if (dict) std::memcpy(dict, b.dict.data(), b.dict.size() * 8);
    return KRB_OK;
}
/* front executor tables: which = 0 piv0, 1 idx_ptr, 2 idx, 3 parent, 4 child_ptr, 5 child, 6 rel, 7 level_ptr, 8 level_fronts,
 * 9 asm_ptr, 10 asm_ent, 11 asm_pos; returns the length (copies when out != NULL) or -1 */
int64_t krb_plan_front_array(const krb_plan *p, int32_t which, int32_t *out) {
    if (!p || !p->p.use_front) return -1;
    const FrontPlan &f = p->p.front;
    const std::vector<int> *v[] = {&f.piv0, &f.idx_ptr, &f.idx, &f.parent, &f.child_ptr, &f.child, &f.rel, &f.level_ptr, &f.level_fronts,
                                   &f.asm_ptr, &f.asm_ent, &f.asm_pos};
    if (which < 0 || which >= 12) return -1;
    if (out && !v[which]->empty()) std::memcpy(out, v[which]->data(), v[which]->size() * sizeof(int));
    return (int64_t)v[which]->size();
}
/* out[0..7] = n_fronts, n_levels, max_m, max_p, lu_size, cb_size, cbv_size, flops */
int32_t krb_plan_front_info(const krb_plan *p, int64_t *out) {
    if (!p || !p->p.use_front || !out) return KRB_ERR_INVALID_ARGUMENT;
    const FrontPlan &f = p->p.front;
    out[0] = f.n_fronts; out[1] = f.n_levels; out[2] = f.max_m; out[3] = f.max_p; out[4] = f.lu_size; out[5] = f.cb_size; out[6] = f.cbv_size; out[7] = f.flops;
    return KRB_OK;
}
// 1: the transversal's static pivot sequence cancelled at the nominal values and was re-pivoted numerically (compiler.cpp static_pivots_ok)
int32_t krb_plan_repivoted(const krb_plan *p) { return p ? (p->p.repivoted ? 1 : 0) : -1; }
int32_t krb_plan_integrator(const krb_plan *p) { return p ? p->p.opts.integrator : -1; }
int32_t krb_plan_state_columns(const krb_plan *p) { return p ? (int32_t)p->p.state_elems.size() : -1; }
int32_t krb_plan_has_step_sections(const krb_plan *p) { return p ? (p->p.has_step_sections ? 1 : 0) : -1; }
// 1: the plan's lane program runs as a circuit-specialised kernel on this device (after its first solve), 0: interpreter / other executor
int32_t krb_plan_jit_active(const krb_plan *p) { return p ? (device_state_has_jit(p->p.dev) ? 1 : 0) : -1; }

static int check_io(const krb_io *io, krb_io &o) {
    krb_default_io(&o);
    if (io) {
        if (io->struct_size == KRB_IO_SIZE_V1) std::memcpy(&o, io, KRB_IO_SIZE_V1);          // callers built against the first layout
        else if (io->struct_size == (int32_t)sizeof(krb_io)) o = *io;
        else { set_error("krb_io.struct_size mismatch"); return KRB_ERR_INVALID_ARGUMENT; }
        o.struct_size = (int32_t)sizeof(krb_io);
    }
    if (o.n_select < 0 || (o.n_select > 0 && !o.select)) { set_error("krb_io: n_select > 0 needs select"); return KRB_ERR_INVALID_ARGUMENT; }
    if (o.n_select > 0 && o.layout != KRB_LAYOUT_POINT_MAJOR) { set_error("krb_io: result selection delivers point-major rows only"); return KRB_ERR_UNSUPPORTED; }
    return KRB_OK;
}
static int set_select(const Plan &P, const krb_io &o, RunArgs &a) {
    a.select = nullptr; a.n_select = 0;
    if (o.n_select <= 0) return KRB_OK;
    for (int k = 0; k < o.n_select; k++)
        if (o.select[k] < 0 || o.select[k] >= P.circuit.n) { set_error("krb_io.select: unknown index out of range"); return KRB_ERR_INVALID_ARGUMENT; }
    a.select = o.select; a.n_select = o.n_select;
    return KRB_OK;
}
static int check_ov(const Plan &P, const krb_overrides *ov, const double **vals) {
    *vals = nullptr;
    const int nv = (int)P.varied.size();
    if (nv == 0) return KRB_OK;
    if (!ov || ov->n_varied != nv || !ov->values) { set_error("plan has varied elements: krb_overrides with matching n_varied is required"); return KRB_ERR_INVALID_ARGUMENT; }
    *vals = ov->values;
    return KRB_OK;
}

int32_t krb_solve_op(krb_plan *plan_dc, const krb_config *cfg, int64_t lanes, const krb_overrides *ov, double *x_out, int32_t *iterations,
                     int32_t *lane_status, const krb_io *io) {
    if (!plan_dc || !cfg || !x_out || lanes < 0 || plan_dc->p.mode != KRB_MODE_DC) { set_error("krb_solve_op: invalid argument"); return KRB_ERR_INVALID_ARGUMENT; }
    if (plan_dc->p.input_sweep >= 0) { set_error("krb_solve_op: plan was analysed with a sweep element; use krb_solve_dc_sweep"); return KRB_ERR_INVALID_ARGUMENT; }
    krb_io o; int rc;
    if ((rc = check_io(io, o))) return rc;
    RunArgs a;
    if ((rc = check_ov(plan_dc->p, ov, &a.varied_values))) return rc;
    a.cfg = *cfg; a.lanes = lanes; a.rows = 1; a.x_out = x_out; a.lane_status = lane_status; a.iterations = iterations;
    a.memory = o.memory; a.layout = o.layout; a.stream = o.stream;
    if ((rc = set_select(plan_dc->p, o, a))) return rc;
    return run_lanes(&plan_dc->p, nullptr, a);
}

int32_t krb_solve_dc_sweep(krb_plan *plan_dc, const krb_config *cfg, const double *sweep_values, int64_t lanes, const krb_overrides *ov,
                           double *x_out, int32_t *lane_status, const krb_io *io) {
    if (!plan_dc || !cfg || !x_out || !sweep_values || lanes < 0 || plan_dc->p.mode != KRB_MODE_DC) { set_error("krb_solve_dc_sweep: invalid argument"); return KRB_ERR_INVALID_ARGUMENT; }
    if (plan_dc->p.input_sweep < 0) { set_error("krb_solve_dc_sweep: plan has no sweep element (krb_plan_opts.sweep_element)"); return KRB_ERR_ELEMENT_NOT_FOUND; }
    krb_io o; int rc;
    if ((rc = check_io(io, o))) return rc;
    RunArgs a;
    if ((rc = check_ov(plan_dc->p, ov, &a.varied_values))) return rc;
    a.cfg = *cfg; a.lanes = lanes; a.rows = 1; a.lane_param = sweep_values; a.x_out = x_out; a.lane_status = lane_status;
    a.memory = o.memory; a.layout = o.layout; a.stream = o.stream;
    if ((rc = set_select(plan_dc->p, o, a))) return rc;
    return run_lanes(&plan_dc->p, nullptr, a);
}

int32_t krb_solve_dc_sweep_sequential(krb_plan *plan_dc, const krb_config *cfg, double start, double step_size, int64_t n_points, int64_t lanes,
                                      const krb_overrides *ov, double *x_out, int32_t *iterations, int32_t *lane_status, const krb_io *io) {
    if (!plan_dc || !cfg || !x_out || lanes < 0 || n_points < 1 || plan_dc->p.mode != KRB_MODE_DC) { set_error("krb_solve_dc_sweep_sequential: invalid argument"); return KRB_ERR_INVALID_ARGUMENT; }
    if (plan_dc->p.input_sweep < 0) { set_error("krb_solve_dc_sweep_sequential: plan has no sweep element (krb_plan_opts.sweep_element)"); return KRB_ERR_ELEMENT_NOT_FOUND; }
    if (n_points > 0x3fffffff) { set_error("krb_solve_dc_sweep_sequential: too many sweep points"); return KRB_ERR_INVALID_ARGUMENT; }
    krb_io o; int rc;
    if ((rc = check_io(io, o))) return rc;
    RunArgs a;
    if ((rc = check_ov(plan_dc->p, ov, &a.varied_values))) return rc;
    a.cfg = *cfg; a.lanes = lanes; a.rows = (int)n_points; a.x_out = x_out; a.lane_status = lane_status; a.iterations = iterations;
    a.sweep_rows = true; a.sweep_start = start; a.sweep_step = step_size;
    a.memory = o.memory; a.layout = o.layout; a.stream = o.stream;
    if ((rc = set_select(plan_dc->p, o, a))) return rc;
    return run_lanes(&plan_dc->p, nullptr, a);
}

int32_t krb_solve_ac(krb_plan *plan_ac, const krb_config *cfg, const double *frequencies, int64_t lanes, const krb_overrides *ov,
                     const double *op_solution, int64_t lanes_op, double *x_out, int32_t *lane_status, const krb_io *io) {
    if (!plan_ac || !cfg || !x_out || !frequencies || lanes < 0 || plan_ac->p.mode != KRB_MODE_AC) { set_error("krb_solve_ac: invalid argument"); return KRB_ERR_INVALID_ARGUMENT; }
    if (plan_ac->p.circuit.nonlinear && !op_solution) { set_error("krb_solve_ac: nonlinear circuit needs op_solution (ac.rs:25)"); return KRB_ERR_INVALID_ARGUMENT; }
    if (op_solution && !(lanes_op == 1 || lanes_op == lanes)) { set_error("krb_solve_ac: lanes_op must be 1 or lanes"); return KRB_ERR_INVALID_ARGUMENT; }
    krb_io o; int rc;
    if ((rc = check_io(io, o))) return rc;
    RunArgs a;
    if ((rc = check_ov(plan_ac->p, ov, &a.varied_values))) return rc;
    a.cfg = *cfg; a.lanes = lanes; a.rows = 1; a.lane_param = frequencies; a.x_out = x_out; a.lane_status = lane_status;
    a.op_solution = plan_ac->p.circuit.nonlinear ? op_solution : nullptr; a.lanes_op = lanes_op;
    a.memory = o.memory; a.layout = o.layout; a.stream = o.stream;
    if ((rc = set_select(plan_ac->p, o, a))) return rc;
    return run_lanes(&plan_ac->p, nullptr, a);
}

int32_t krb_solve_transient(krb_plan *plan_dc, krb_plan *plan_tran, const krb_config *cfg, int64_t n_steps, int64_t lanes,
                            const krb_overrides *ov, double *x_out, int32_t *lane_status, const krb_io *io) {
    if (!plan_dc || !plan_tran || !cfg || !x_out || lanes < 0 || n_steps < 0 || plan_dc->p.mode != KRB_MODE_DC || plan_tran->p.mode != KRB_MODE_TRAN ||
        plan_dc->p.circuit.n != plan_tran->p.circuit.n) { set_error("krb_solve_transient: invalid argument"); return KRB_ERR_INVALID_ARGUMENT; }
    if (plan_dc->p.input_sweep >= 0) { set_error("krb_solve_transient: plan_dc must not have a sweep element"); return KRB_ERR_INVALID_ARGUMENT; }
    if (n_steps > 0x3fffffff) { set_error("krb_solve_transient: too many steps"); return KRB_ERR_INVALID_ARGUMENT; }
    krb_io o; int rc;
    if ((rc = check_io(io, o))) return rc;
    RunArgs a;
    if ((rc = check_ov(plan_dc->p, ov, &a.varied_values))) return rc;
    const double *v2;
    if ((rc = check_ov(plan_tran->p, ov, &v2))) return rc;
    a.cfg = *cfg; a.lanes = lanes; a.rows = (int)n_steps + 1; a.x_out = x_out; a.lane_status = lane_status;
    a.memory = o.memory; a.layout = o.layout; a.stream = o.stream;
    if ((rc = set_select(plan_dc->p, o, a))) return rc;
    return run_lanes(&plan_dc->p, &plan_tran->p, a);
}

const char *krb_last_error_string(void) { return get_error(); }
const char *krb_version(void) { return "krets_b200 0.2 (sm_100a: chain, lane-VM, dense-transient and front executors)"; }
int32_t krb_device_count(void) { return device_count(); }
int64_t krb_kernel_launch_count(void) { return launch_count(); }
double krb_last_kernel_ms(void) { return last_kernel_ms(); }

}  // extern "C"
