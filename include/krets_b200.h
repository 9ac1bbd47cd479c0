/*
 * krets_b200 — C ABI of the B200-native batched solver backend for prokie/krets.
 *
 * This header is the drop-in boundary for the krets-solver hot path.  The reference has no FFI seam; its
 * seam is the Rust API
 *     Solver::new(Circuit, SolverConfig)            crates/krets-solver/src/solver.rs:23
 *     Solver::solve(&mut self, Analysis)            crates/krets-solver/src/solver.rs:31-52
 * i.e. the four free functions solver::{op,dc,ac,transient}::solve (op.rs:15, dc.rs:14, ac.rs:17,
 * transient.rs:14).  A thin Rust shim flattens `Circuit` into the POD element table below
 * (index_map lookups done once on the host, exactly parser.rs:105-125) and forwards to these entry points.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, int32 status return (KRB_OK == 0), never throws/aborts.
 *   - status codes map 1:1 onto krets_solver::Error variants (error.rs:4-36); per-lane conditions that the
 *     reference would turn into an early `Err` are reported in `lane_status[]` instead.
 *   - a "lane" is one independent solve: a DC sweep value, an AC frequency or a Monte-Carlo variant.
 *   - outputs: caller-allocated; `krb_io.memory` says whether pointers are host or device memory.
 *   - the library owns plans and device scratch; a plan is used by one thread at a time; all kernels of a
 *     call are enqueued on `krb_io.stream` (a cudaStream_t; NULL = default stream).
 *   - there is NO CPU fallback: every numeric entry point fails with KRB_ERR_NO_DEVICE without a CUDA device.
 */
#ifndef KRETS_B200_H
#define KRETS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------------------------------------- */
enum {
    KRB_OK = 0,
    KRB_ERR_INVALID_ARGUMENT = 1,
    KRB_ERR_ELEMENT_NOT_FOUND = 3,      /* Error::ElementNotFound          error.rs:10-12, dc.rs:24-28 */
    KRB_ERR_INVALID_ELEMENT_FORMAT = 4, /* Error::InvalidElementFormat     error.rs:18-20, dc.rs:34-38 */
    KRB_ERR_MAX_ITERATIONS = 5,         /* Error::MaximumIterationsExceeded error.rs:14-16, op.rs:79-81 */
    KRB_ERR_DECOMPOSITION_FAILED = 6,   /* Error::DecompositionFailed / MatrixDecomposition  error.rs:26-36 */
    KRB_ERR_UNSUPPORTED = 7,            /* reference panics here (todo!(): stampable.rs:327,336,356,920) */
    KRB_ERR_NO_DEVICE = 20,
    KRB_ERR_CUDA = 21,
    KRB_ERR_OUT_OF_MEMORY = 22,
    KRB_ERR_INTERNAL = 23
};

/* per-lane status bits (lane_status[]) */
enum {
    KRB_LANE_OK = 0,
    KRB_LANE_ZERO_PIVOT = 1,            /* a static pivot was exactly zero or non-finite (-> DecompositionFailed) */
    KRB_LANE_NOT_CONVERGED = 2,         /* NR hit maximum_iterations (-> MaximumIterationsExceeded) */
    KRB_LANE_NONFINITE = 4,             /* a delivered unknown is inf/nan */
    KRB_LANE_INACCURATE = 8             /* pivot growth: a scaled factor entry (u^ = u/pivot, or a front's multiplier l/pivot) reached
                                           2^24 on this lane, i.e. the STATIC pivot cancelled >= 24 bits against its row — where the
                                           reference's faer LU (op.rs:50, ac.rs:70) would have exchanged rows.  The lane's unknowns may
                                           have lost >= 7 of 16 digits (error ~ eps * growth >= 1e-9, the parity contract).  The host
                                           mirror re-solves such lanes on the front executor (pivoting inside fronts + refinement).
                                           Linear circuits only: over Newton iterates a diode conductance spans tens of decades, so the
                                           measure says nothing there and the convergence check (solver.rs:76-104) is the guard */
};
#define KRB_GROWTH_LIMIT_LOG2 24

/* ---- circuit (flattened krets_parser::circuit::Circuit, circuit.rs:7-20) ------------------------------ */
/* element kinds; same order as the reference's parser `alt`, elements.rs:45-55 */
enum { KRB_R = 0, KRB_C = 1, KRB_L = 2, KRB_V = 3, KRB_I = 4, KRB_D = 5, KRB_NMOS = 6 };
enum { KRB_WAVE_NONE = 0, KRB_WAVE_PULSE = 1, KRB_WAVE_SIN = 2 };
enum { KRB_ELEM_FLAG_G2 = 1,            /* capacitor with the G2 token (capacitor.rs:31,38) */
       /* EXTENSION (not reference behaviour): this NMOS stamps the level-1 companion model the reference's
        * nmosfet.rs:51-102 equations define (g_m, g_ds, i_d; documentation/main.typ:375-424) with KCL-consistent signs.
        * Without the flag an NMOS reproduces the reference: its stamps look up bare node names, never match the index map
        * and stamp nothing (stampable.rs:842-844, circuits/inverter/dc/result.parquet). */
       KRB_ELEM_FLAG_LEVEL1 = 2 };

typedef struct krb_element {
    int32_t kind;
    int32_t ip;            /* index_map["V(<plus>)"]   or -1 (ground / absent)   NMOS: drain  */
    int32_t im;            /* index_map["V(<minus>)"]  or -1                     NMOS: gate   */
    int32_t ic;            /* index_map["I(<ident>)"]  or -1 (group-1 element)   NMOS: source */
    int32_t waveform;      /* KRB_WAVE_* for V sources (voltage_source.rs:244-252) */
    int32_t flags;
    double value;          /* R/C/L value, V dc_value, I value */
    double ac_amplitude;   /* V sources */
    double wave[7];        /* PULSE: v0 v1 td tr tf pw per (voltage_source.rs:15-31); SIN: off amp freq delay damp phase;
                              NMOS: Vth, beta = KP*(W/L), lambda (models/nmosfet.rs:24-45) */
    double is;             /* diode saturation current   (models/diode.rs:14-23) */
    double n;              /* diode emission coefficient */
} krb_element;

typedef struct krb_config {          /* krets_solver::config::SolverConfig, config.rs:4-37 */
    double relative_tolerance;
    double current_absolute_tolerance;
    double voltage_absolute_tolerance;
    int32_t maximum_iterations;
} krb_config;

typedef struct krb_circuit krb_circuit;
typedef struct krb_plan krb_plan;

void krb_default_config(krb_config *cfg);                              /* config.rs:26-36 */

/* elements in netlist order; n_unknowns = index_map.len(); unknown_is_current[i] != 0 iff the i-th key starts
 * with 'I' (solver.rs:96-100).  The table is copied. */
int32_t krb_circuit_create(const krb_element *elements, int32_t n_elements, int32_t n_unknowns,
                           const uint8_t *unknown_is_current, krb_circuit **out);
void krb_circuit_destroy(krb_circuit *c);

/* ---- one-time symbolic analysis ----------------------------------------------------------------------- */
enum { KRB_MODE_DC = 0, KRB_MODE_AC = 1, KRB_MODE_TRAN = 2 };          /* stamp_*_dc / _ac / _transient */
enum { KRB_INTEGRATOR_BE_KRETS = 0,   /* backward Euler incl. the reference's one-step source lag */
       KRB_INTEGRATOR_TRAP = 1,       /* trapezoidal companion models, sources at t_n (north-star extension; not in the reference) */
       KRB_INTEGRATOR_BE = 2 };       /* backward Euler with sources evaluated at t_n (no lag; not reference behaviour) */
enum { KRB_EXEC_FUSED = 0,            /* default: stamp+factor+solve fused per lane; the executor is chosen from the plan's
                                         structure (native chain kernel for tridiagonal-after-permutation plans, lane VM otherwise) */
       KRB_EXEC_STAGED = 1,           /* front executor: stamp, supernodal multifrontal factor and level-scheduled solves as separate
                                         launches over per-lane factor stores in HBM (meshes; linear circuits).  Also chosen by
                                         KRB_EXEC_FUSED when the fill is too large for a per-lane program */
       KRB_EXEC_LANE_VM = 2 };        /* force the general lane-program interpreter */

typedef struct krb_plan_opts {
    int32_t struct_size;              /* sizeof(krb_plan_opts), for ABI evolution */
    int32_t exec;                     /* KRB_EXEC_* */
    int32_t integrator;               /* KRB_INTEGRATOR_* (TRAN) */
    int32_t scratch_slots;            /* per-lane on-chip scratch capacity (0 = default) */
    double time_step;                 /* h (TRAN) */
    double nominal_frequency;         /* AC: frequency at which the static pivot plan is chosen (0 = 1 kHz); for the front
                                         executor use the LOWEST frequency of the sweep (smallest |jwL| branch diagonals) */
    int32_t sweep_element;            /* DC: element ordinal whose value is the per-lane sweep value, or -1 */
    int32_t n_varied;                 /* Monte-Carlo: number of elements whose value is a per-lane input */
    const int32_t *varied_elements;   /* [n_varied] element ordinals (R/C/L/V/I) */
    int32_t threads_per_block;        /* 0 = default */
    int32_t refine_steps;             /* front executor: 0 = default (AC plans: 2 steps of iterative refinement on the stamped matrix,
                                         DC/TRAN: none); k > 0 = k steps, followed by a componentwise backward-error check of every
                                         delivered row (|b - A x| <= 2^-40 (|A||x| + |b|)) that sets or CLEARS KRB_LANE_INACCURATE —
                                         the a-posteriori answer replaces the a-priori growth flag.  Used by the host mirror's
                                         pivoting fallback */
} krb_plan_opts;

void krb_default_plan_opts(krb_plan_opts *o);

/* Host-only: stamp pattern, static pivot plan (weighted matching + minimum degree), symbolic LU, lane program.
 * KRB_ERR_DECOMPOSITION_FAILED if the pattern is structurally singular (the reference's faer error, e.g. AC with
 * a current source); KRB_ERR_UNSUPPORTED for NMOS in AC (stampable.rs:914-921). */
int32_t krb_analyze(const krb_circuit *c, int32_t mode, const krb_plan_opts *opts, krb_plan **out);
void krb_plan_destroy(krb_plan *p);

typedef struct krb_plan_info {
    int32_t n;                /* unknowns */
    int32_t nnz;              /* unique stamped (row,col) = nnz of the reference's CSC */
    int32_t nnz_lu;           /* nnz(L+U) of the static plan (strict L + U incl. diagonal) */
    int32_t mode;
    int32_t nonlinear;        /* circuit has a diode/NMOS (elements.rs:143-148) */
    int32_t scratch_slots;    /* on-chip scratch cells per lane actually used */
    int32_t spill_planes;     /* HBM value planes per lane used by the program */
    int32_t n_inputs;         /* per-lane input planes (sweep value / frequency / varied elements) */
    int64_t program_ops;      /* lane-program length */
    int64_t flops_per_point;  /* real flops of factor+solve for one lane (complex counted x4/x2) */
    int64_t hbm_bytes_per_point;   /* bytes the lane program itself moves through HBM per point (model) */
    int64_t alg_bytes_per_point;   /* P + s*(4*nnz_lu + 3*n) */
    int32_t executor;              /* 0 lane VM, 1 chain kernel, 2 front executor */
    int32_t reserved;
} krb_plan_info;

int32_t krb_plan_get_info(const krb_plan *p, krb_plan_info *info);
/* CSC pattern of the stamped matrix (sorted rows) — must be bit-identical to the reference's after sum_triplets */
int32_t krb_plan_get_pattern(const krb_plan *p, int32_t *colptr /*n+1*/, int32_t *rowidx /*nnz*/);
/* static pivot plan: pivot k eliminates unknown col_order[k] using equation row_order[k] */
int32_t krb_plan_get_pivot_order(const krb_plan *p, int32_t *row_order /*n*/, int32_t *col_order /*n*/);
/* lane program introspection (tests run it through a numpy emulator on CPU-only machines):
 * section 0 = first/only section, 1 = factor-once section, 2 = per-step section.  ops are 2 x uint32 each. */
int64_t krb_plan_program_size(const krb_plan *p, int32_t section);
int32_t krb_plan_get_program(const krb_plan *p, int32_t section, uint32_t *ops /*2*size*/);
/* Circuit-specialised lane kernel.  Lane programs of up to KRB_VM_JIT_MAX_OPS (default 20000) operations are not interpreted: at
 * the plan's first solve (always up to 2500 operations; above, when that solve has at least KRB_VM_JIT_MIN_POINTS = 32768 lanes x rows to
 * pay for NVRTC's half second per thousand operations) the library writes the program out as straight-line CUDA (one statement per operation, scratch cells in registers),
 * compiles it with NVRTC for sm_100a and launches that kernel (KRB_VM_JIT=0: always interpret).  The two entry points below expose
 * the generated text and the NVRTC step without a device.  krb_plan_jit_source: returns the size of the text incl. the terminating
 * NUL (0: plan not eligible, -1: bad arguments) and copies at most cap bytes; host_mode != 0 prefixes the text for a host build
 * behind tests/jit_host_shim.h (the CPU tests run the generated statements against the oracle that way). */
int64_t krb_plan_jit_source(const krb_plan *p, int32_t host_mode, char *buf, int64_t cap);
int32_t krb_plan_jit_compile(const krb_plan *p, int64_t *cubin_bytes /* may be NULL */);

/* ---- I/O descriptors -------------------------------------------------------------------------------- */
enum { KRB_MEM_HOST = 0, KRB_MEM_DEVICE = 1 };
enum { KRB_LAYOUT_POINT_MAJOR = 0,    /* x[lane][row][unknown]  — one contiguous vector per solved point (reference rows) */
       KRB_LAYOUT_LANE_INTERLEAVED = 1 /* x[row][unknown][lane] — the kernels' native layout */ };

typedef struct krb_io {
    int32_t struct_size;
    int32_t memory;                   /* KRB_MEM_*: applies to every lane-sized pointer of the call */
    int32_t layout;                   /* KRB_LAYOUT_* of x_out */
    int32_t reserved;
    void *stream;                     /* cudaStream_t */
    /* EXTENSION (not reference behaviour: ac.rs:97-105, dc.rs:92-97 and transient.rs:90-95 return every unknown of every point):
     * result selection.  n_select > 0: only the unknowns select[0..n_select) (index-map indices, HOST pointer, copied by the call)
     * are delivered and x_out holds n_select values per solved point instead of n (point-major layout only).  What a plotting
     * front end (krets-gui) needs; cuts the device->host traffic of a sweep from n to n_select values per point. */
    const int32_t *select;
    int32_t n_select;
    int32_t reserved2;
} krb_io;
#define KRB_IO_SIZE_V1 24             /* struct_size of the layout without the selection fields (still accepted) */
void krb_default_io(krb_io *io);

/* per-lane element values (Monte-Carlo): values[k][lane] is the value of plan.varied_elements[k] on that lane */
typedef struct krb_overrides {
    int32_t n_varied;                 /* must equal the plan's n_varied (0 allowed with values == NULL) */
    int32_t reserved;
    const double *values;             /* [n_varied][lanes] */
} krb_overrides;

/* ---- analyses ------------------------------------------------------------------------------------------ */
/* op.rs:15-86 on `lanes` Monte-Carlo variants (lanes == 1 and ov == NULL is the reference call).
 * x_out: lanes*n doubles.  iterations (optional, [lanes]): NR iterations used. */
int32_t krb_solve_op(krb_plan *plan_dc, const krb_config *cfg, int64_t lanes, const krb_overrides *ov, double *x_out,
                     int32_t *iterations, int32_t *lane_status, const krb_io *io);

/* dc.rs:14-124 with the sweep points as parallel lanes: lane l solves with the plan's sweep element set to sweep_values[l]
 * (= start + l*step on the host).  Every lane starts Newton-Raphson from the empty guess, so for NONLINEAR circuits this is not the
 * reference's iteration history (dc.rs:58 warm start) — use krb_solve_dc_sweep_sequential for that; for linear circuits the two agree. */
int32_t krb_solve_dc_sweep(krb_plan *plan_dc, const krb_config *cfg, const double *sweep_values, int64_t lanes,
                           const krb_overrides *ov, double *x_out, int32_t *lane_status, const krb_io *io);

/* dc.rs:47-113 as written, for nonlinear circuits: the n_points sweep points are the result ROWS of a lane and are solved in order, the
 * sweep element at start + r*step_size (dc.rs:48); row 0 starts Newton-Raphson from the empty map (dc.rs:41), row r from row r-1's
 * solution (dc.rs:58,112), and the "step" key of the result map (dc.rs:97) takes part in the convergence check (solver.rs:86-103).
 * The parallel dimension is the Monte-Carlo lane (lanes == 1 and ov == NULL is the reference call).
 * x_out: lanes*n_points*n doubles, x[lane][row][unknown] (krb_io.layout POINT_MAJOR) — the layout of krb_solve_transient.
 * iterations (optional, [lanes]): NR iterations of the last row.  Plans that do not run on the lane VM (linear chain / front plans,
 * where the warm start cannot matter) return KRB_ERR_UNSUPPORTED. */
int32_t krb_solve_dc_sweep_sequential(krb_plan *plan_dc, const krb_config *cfg, double start, double step_size, int64_t n_points,
                                      int64_t lanes, const krb_overrides *ov, double *x_out, int32_t *iterations, int32_t *lane_status,
                                      const krb_io *io);

/* ac.rs:39-105: lane l solves at frequencies[l] (> 0; the host drops f <= 0 as ac.rs:40-45).
 * op_solution: [n] (lanes_op == 1) or [lanes][n] point-major DC operating point, needed only for nonlinear
 * circuits (diode small-signal conductance, stampable.rs:614-641); may be NULL for linear circuits.
 * x_out: complex128 (re,im) pairs, lanes*n of them. */
int32_t krb_solve_ac(krb_plan *plan_ac, const krb_config *cfg, const double *frequencies, int64_t lanes,
                     const krb_overrides *ov, const double *op_solution, int64_t lanes_op, double *x_out,
                     int32_t *lane_status, const krb_io *io);

/* transient.rs:14-114: OP (row 0, through plan_dc) then n_steps fixed steps of plan_tran.time_step.
 * x_out: lanes*(n_steps+1)*n doubles; time of row r is r*h. */
int32_t krb_solve_transient(krb_plan *plan_dc, krb_plan *plan_tran, const krb_config *cfg, int64_t n_steps, int64_t lanes,
                            const krb_overrides *ov, double *x_out, int32_t *lane_status, const krb_io *io);

/* ---- misc -------------------------------------------------------------------------------------------- */
const char *krb_last_error_string(void);        /* thread-local, valid until the next call on this thread */
const char *krb_version(void);
int32_t krb_device_count(void);                 /* 0 without a usable CUDA device */
/* counters for bench.py: kernels launched by this library since process start */
int64_t krb_kernel_launch_count(void);
/* device time of the most recent call's lane-program kernel(s), in ms (CUDA events on io.stream); <0 if none */
double krb_last_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* KRETS_B200_H */
