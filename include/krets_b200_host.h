/*
 * krets_b200 host mirror — the krets-parser / krets-solver public API restated over the C ABI of krets_b200.h.
 *
 * The reference's host is Rust; this image has no Rust toolchain, so the same surface is provided in C++ (compiled
 * into libkrets_b200.so) and exported through these C handles so that tests and tools can drive it:
 *     parse_circuit_description{,_file}   crates/krets-parser/src/parser.rs:27,156
 *     Solver::new / Solver::solve         crates/krets-solver/src/solver.rs:23,31-52
 *     AnalysisResult                      crates/krets-solver/src/lib.rs:13-30
 *     AcAnalysis::generate_frequencies    crates/krets-parser/src/analyses.rs:80-158
 * Everything numeric goes through krb_solve_* (CUDA); parsing, index assignment and frequency grids are host work.
 */
#ifndef KRETS_B200_HOST_H
#define KRETS_B200_HOST_H

#include "krets_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* parser-side error codes (krets_parser::Error, crates/krets-parser/src/error.rs) */
enum {
    KRH_ERR_PARSE = 30,           /* ParseError / InvalidFormat */
    KRH_ERR_EMPTY_NETLIST = 31,   /* EmptyNetlist */
    KRH_ERR_UNSUPPORTED = 32,     /* BJT / subcircuit instance lines (outside the restated subset) */
    KRH_ERR_MISSING_MODEL = 33,   /* reference: todo!() parser.rs:138,146 */
    KRH_ERR_IO = 34
};

typedef struct krh_circuit krh_circuit;
typedef struct krh_solver krh_solver;
typedef struct krh_result krh_result;

/* ---- krets_parser::parser ---- */
int32_t krh_parse_circuit_description(const char *text, krh_circuit **out);
int32_t krh_parse_circuit_description_file(const char *path, krh_circuit **out);
void krh_circuit_free(krh_circuit *c);
int32_t krh_circuit_num_unknowns(const krh_circuit *c);           /* index_map.len() */
const char *krh_circuit_unknown_name(const krh_circuit *c, int32_t index);
int32_t krh_circuit_index_of(const krh_circuit *c, const char *key);   /* index_map.get(key), -1 if absent */
int32_t krh_circuit_num_elements(const krh_circuit *c);
const char *krh_circuit_element_identifier(const krh_circuit *c, int32_t e);
const krb_element *krh_circuit_elements(const krh_circuit *c);    /* the flattened table krb_circuit_create takes */
const uint8_t *krh_circuit_is_current(const krh_circuit *c);
/* EXTENSION (not reference behaviour).  models = KRH_DEVICES_REFERENCE (default): an NMOS reproduces the reference and
 * stamps nothing (stampable.rs:842-844).  KRH_DEVICES_LEVEL1: every NMOS of the circuit gets KRB_ELEM_FLAG_LEVEL1 and stamps the
 * level-1 companion model of nmosfet.rs:51-102 (OP / DC / transient: Newton-Raphson; AC: small-signal g_m, g_ds at the operating
 * point).  Call before the circuit is handed to krh_solver_new / krb_circuit_create.  Returns the number of NMOS elements, < 0 on error. */
enum { KRH_DEVICES_REFERENCE = 0, KRH_DEVICES_LEVEL1 = 1 };
int32_t krh_circuit_set_device_models(krh_circuit *c, int32_t models);

/* ---- krets_parser::analyses::AcAnalysis::generate_frequencies ---- */
enum { KRH_SWEEP_LINEAR = 0, KRH_SWEEP_DECADE = 1, KRH_SWEEP_OCTAVE = 2 };
int64_t krh_generate_frequencies(int32_t kind, uint32_t points, double fstart, double fstop, double *out, int64_t cap);
int32_t krh_dc_num_steps(double start, double stop, double step_size);       /* dc.rs:44-45 */
int32_t krh_transient_num_steps(double time_step, double stop_time);         /* transient.rs:29 */

/* ---- krets_solver::solver::Solver ---- */
int32_t krh_solver_new(const krh_circuit *c, const krb_config *cfg, krh_solver **out);
void krh_solver_free(krh_solver *s);
int32_t krh_solver_solve_op(krh_solver *s, krh_result **out);                                        /* Analysis::Op */
/* Analysis::Dc, dc.rs:14-124.  Linear circuits: the sweep points are parallel lanes.  Nonlinear circuits: the points are solved in order,
 * Newton-Raphson warm-started from the previous point exactly as dc.rs:58,112 (krb_solve_dc_sweep_sequential). */
int32_t krh_solver_solve_dc(krh_solver *s, const char *element, double start, double stop, double step_size, krh_result **out);
/* extension (not in the reference): KRH_DC_PARALLEL_COLD_START solves the points of a NONLINEAR sweep as parallel lanes, each from the
 * empty guess — the fast path when the sweep is long; results agree with the reference's within the Newton-Raphson tolerances, not to 1e-9 */
enum { KRH_DC_REFERENCE = 0, KRH_DC_PARALLEL_COLD_START = 1 };
int32_t krh_solver_solve_dc_with(krh_solver *s, const char *element, double start, double stop, double step_size, int32_t mode, krh_result **out);
int32_t krh_solver_solve_ac(krh_solver *s, int32_t sweep_kind, uint32_t points, double fstart, double fstop, krh_result **out);
int32_t krh_solver_solve_transient(krh_solver *s, double time_step, double stop_time, krh_result **out);
/* extension (not in the reference): same driver with KRB_INTEGRATOR_TRAP / KRB_INTEGRATOR_BE instead of the reference's lagged backward Euler */
int32_t krh_solver_solve_transient_with(krh_solver *s, double time_step, double stop_time, int32_t integrator, krh_result **out);

/* ---- krets_solver::AnalysisResult ---- */
enum { KRH_RESULT_OP = 0, KRH_RESULT_DC = 1, KRH_RESULT_AC = 2, KRH_RESULT_TRANSIENT = 3 };
int32_t krh_result_kind(const krh_result *r);
int64_t krh_result_rows(const krh_result *r);
int32_t krh_result_num_unknowns(const krh_result *r);
const double *krh_result_data(const krh_result *r);   /* [rows][n] doubles (AC: [rows][n][2] re,im) in index-map order */
const double *krh_result_axis(const krh_result *r);   /* per row: "step" (DC), "frequency" (AC), "time" (transient); NULL for OP */
void krh_result_free(krh_result *r);

const char *krh_last_error_string(void);

#ifdef __cplusplus
}
#endif
#endif
