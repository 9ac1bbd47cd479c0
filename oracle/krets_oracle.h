/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the prokie/krets solver hot path (stamp -> sum_triplets -> CSC -> sparse LU ->
 * solve, wrapped by the OP / DC / AC / transient drivers).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library, and only as the checker or
 * the reported CPU baseline — never as the product.  Every function cites the reference file:line
 * it follows (paths relative to /root/reference).
 *
 * Parity status: stamps, index map, frequency grids, waveforms, NR/convergence logic follow in-tree
 * reference code line by line and are pinned by the reference's own test assertions and the
 * full-precision golden circuits/low_pass_filter/ac/result.parquet (tests/golden/).  The sparse LU
 * lives in the un-vendored crate faer 0.23.2; its bit-level behaviour is "parity unpinned" (see
 * ko_lu_template.h).
 */
#ifndef KRETS_ORACLE_H
#define KRETS_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct ko_circuit ko_circuit;

/* error codes: 1:1 with krets_solver::Error (crates/krets-solver/src/error.rs:4-36) and the parser's */
enum {
    KO_OK = 0,
    KO_ERR_PARSE = 1,                 /* krets_parser::Error::ParseError / InvalidFormat */
    KO_ERR_EMPTY_NETLIST = 2,         /* krets_parser::Error::EmptyNetlist (parser.rs:127-129) */
    KO_ERR_ELEMENT_NOT_FOUND = 3,     /* Error::ElementNotFound (dc.rs:24-28) */
    KO_ERR_INVALID_ELEMENT_FORMAT = 4,/* Error::InvalidElementFormat (dc.rs:34-38) */
    KO_ERR_MAX_ITERATIONS = 5,        /* Error::MaximumIterationsExceeded (op.rs:79-81) */
    KO_ERR_DECOMPOSITION = 6,         /* Error::DecompositionFailed / MatrixDecomposition */
    KO_ERR_UNSUPPORTED = 7,           /* reference panics (todo!()) or element outside the restated subset */
    KO_ERR_BUFFER = 8,
    KO_ERR_MISSING_MODEL = 9          /* reference: todo!() at parser.rs:138,146 */
};

/* element kinds, in the order of the reference's `alt` in elements.rs:45-55 */
enum { KO_R = 0, KO_C = 1, KO_L = 2, KO_V = 3, KO_I = 4, KO_D = 5, KO_NMOS = 6 };

typedef struct {
    double relative_tolerance;          /* config.rs:29 */
    double current_absolute_tolerance;  /* config.rs:30 */
    double voltage_absolute_tolerance;  /* config.rs:31 */
    int maximum_iterations;             /* config.rs:32 */
} ko_config;

void ko_default_config(ko_config *cfg);

/* parser.rs:27-154 (subset R/C/L/V/I/D/MN + .model D/NMOS + comment/.control/.subckt skipping) */
ko_circuit *ko_parse(const char *text, int *err, char *errmsg, int errmsg_len);
void ko_free(ko_circuit *c);
int ko_num_unknowns(const ko_circuit *c);
int ko_num_elements(const ko_circuit *c);
const char *ko_unknown_name(const ko_circuit *c, int idx);      /* "V(in)" / "I(V1)", index-map order */
int ko_index_of(const ko_circuit *c, const char *name);         /* -1 if absent */
int ko_element_index(const ko_circuit *c, const char *ident);   /* by identifier, e.g. "R1"; -1 if absent */
const char *ko_element_ident(const ko_circuit *c, int e);
int ko_element_kind(const ko_circuit *c, int e);
double ko_element_value(const ko_circuit *c, int e);            /* R/C/L value, V dc_value, I value */
void ko_set_element_value(ko_circuit *c, int e, double v);
/* flattened element table: ip/im/ic = index of V(plus)/V(minus)/I(ident), -1 if absent */
void ko_element_indices(const ko_circuit *c, int e, int *ip, int *im, int *ic);
int ko_is_nonlinear(const ko_circuit *c);

/* Stamped, summed, CSC-sorted matrix for one mode (pattern checks).  mode: 0 DC/OP, 1 AC, 2 transient.
 * vals has 1 (real) or 2 (re,im) doubles per entry.  `guess` may be NULL (empty solution map).
 * returns nnz, or -needed if cap is too small. */
int ko_stamp_matrix(const ko_circuit *c, int mode, double freq_or_h, const double *guess, int *colptr /*n+1*/, int *rowidx,
                    double *vals, int cap);

/* analyses.rs:80-158; kind: 0 Linear{total_points}, 1 Decade{points_per_decade}, 2 Octave{points_per_octave} */
int ko_generate_frequencies(int kind, unsigned points, double fstart, double fstop, double *out, int cap);

/* voltage_source.rs:244-252 for element e (must be a V source) */
double ko_transient_value_at(const ko_circuit *c, int e, double time);

/* op.rs:15-86.  x[n] in index-map order. */
int ko_solve_op(const ko_circuit *c, const ko_config *cfg, double *x, int *iterations);

/* dc.rs:14-124.  out[npts][n]; warm_start=1 is the reference (sequential warm start dc.rs:58,112);
 * warm_start=0 starts every point from the empty map (NOT reference behaviour; used to check the
 * lane-parallel product, identical for linear circuits).  nthreads>1 splits points over threads
 * (only valid when results do not depend on the warm start, i.e. linear circuits or warm_start=0). */
int ko_dc_num_steps(double start, double stop, double step);
/* SURVEY 8(d) cpu_best_effort: analysis once, per point re-stamp + numeric LU + solve, OpenMP over points; linear circuits */
int ko_solve_dc_best_effort(ko_circuit *c, const char *element, double start, double step, int num, double *out, int nthreads);
int ko_solve_dc(ko_circuit *c, const ko_config *cfg, const char *element, double start, double stop, double step,
                double *out, int cap_points, int *npoints, int warm_start, int nthreads);

/* ac.rs:17-109.  out[nf][n][2]; frequencies <= 0 are skipped exactly as ac.rs:40-45 (kept[] tells which). */
int ko_solve_ac(const ko_circuit *c, const ko_config *cfg, const double *freqs, int nf, double *out, int *nout, int nthreads);

/* transient.rs:14-114.  out[(nsteps+1)][n] rows; time[row] = row*h (row 0 = OP). */
int ko_tran_num_steps(double h, double stop);
int ko_solve_tran(const ko_circuit *c, const ko_config *cfg, double h, double stop, double *out, double *time,
                  int cap_rows, int *nrows);

/* integrator: 0 reference BE (== ko_solve_tran), 1 trapezoidal, 2 BE without the source lag — 1 and 2 are not reference behaviour */
int ko_solve_tran_ex(const ko_circuit *c, const ko_config *cfg, double h, double stop, int integrator, double *out, double *time,
                     int cap_rows, int *nrows);

/* Bounded faithful-work samples for bench.py's cpu_baseline: solve `npoints` DC points / AC points of the
 * circuit exactly as the drivers above do (full re-stamp, hash-sum, CSC, symbolic+numeric LU, solve per point). */
const char *ko_version(void);

#ifdef __cplusplus
}
#endif
#endif
