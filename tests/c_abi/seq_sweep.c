/*
 * Plain-C client of the sequential DC sweep (dc.rs:47-113 as written: nonlinear sweep points solved in order, Newton-Raphson
 * warm-started from the previous point, dc.rs:58,112).  No Python, no torch: starts in milliseconds on a GPU box.
 *
 *   seq_sweep <netlist> <element> <start> <stop> <step> [<varied element> <value> <value> ...]
 *
 * prints one JSON object:
 *   "reference":  krh_solver_solve_dc            (Solver::solve(Analysis::Dc): sequential for nonlinear circuits)        [rows][n]
 *   "parallel":   krh_solver_solve_dc_with(KRH_DC_PARALLEL_COLD_START)  (extension: the points as parallel lanes)        [rows][n]
 *   "lanes":      krb_solve_dc_sweep_sequential with one Monte-Carlo lane per <value> of <varied element>                 [lanes][rows][n]
 *   "iterations": NR iterations of the last row of every lane, "status": lane status words
 * tests/test_c_abi.py compares them with the oracle (warm_start = 1 / 0).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "krets_b200_host.h"

#define CHECK(cond, msg)                                                                     \
    do {                                                                                     \
        if (!(cond)) { fprintf(stderr, "seq_sweep: %s (%s)\n", msg, krb_last_error_string()); return 1; } \
    } while (0)

static void print_rows(const char *key, const double *x, long rows, int n, int last) {
    printf(" \"%s\": [", key);
    for (long r = 0; r < rows; r++) {
        printf("%s[", r ? ", " : "");
        for (int u = 0; u < n; u++) printf("%s%.17g", u ? ", " : "", x[r * n + u]);
        printf("]");
    }
    printf("]%s\n", last ? "" : ",");
}

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: seq_sweep <netlist> <element> <start> <stop> <step> [<varied element> <values...>]\n"); return 2; }
    const char *element = argv[2];
    const double start = atof(argv[3]), stop = atof(argv[4]), step = atof(argv[5]);
    krh_circuit *c = NULL;
    CHECK(krh_parse_circuit_description_file(argv[1], &c) == KRB_OK && c, "krh_parse_circuit_description_file");
    const int n = krh_circuit_num_unknowns(c);
    const int num = krh_dc_num_steps(start, stop, step);
    krb_config cfg;
    krb_default_config(&cfg);
    krh_solver *s = NULL;
    CHECK(krh_solver_new(c, &cfg, &s) == KRB_OK && s, "krh_solver_new");
    if (krb_device_count() <= 0) {
        krh_result *r = NULL;
        CHECK(krh_solver_solve_dc(s, element, start, stop, step, &r) == KRB_ERR_NO_DEVICE, "without a GPU the sweep must fail with KRB_ERR_NO_DEVICE");
        printf("{\"no_device\": true}\n");
        return 0;
    }
    printf("{\n \"n\": %d, \"rows\": %d,\n", n, num);
    krh_result *r = NULL;
    CHECK(krh_solver_solve_dc(s, element, start, stop, step, &r) == KRB_OK && r && krh_result_rows(r) == num, "krh_solver_solve_dc");
    print_rows("reference", krh_result_data(r), num, n, 0);
    krh_result_free(r);
    r = NULL;
    CHECK(krh_solver_solve_dc_with(s, element, start, stop, step, KRH_DC_PARALLEL_COLD_START, &r) == KRB_OK && r, "krh_solver_solve_dc_with");
    print_rows("parallel", krh_result_data(r), num, n, argc <= 7);
    krh_result_free(r);
    if (argc > 7) {
        /* Monte-Carlo lanes through the C ABI itself */
        const int lanes = argc - 7;
        int ei = -1, vi = -1;
        for (int e = 0; e < krh_circuit_num_elements(c); e++) {
            if (!strcmp(krh_circuit_element_identifier(c, e), element)) ei = e;
            if (!strcmp(krh_circuit_element_identifier(c, e), argv[6])) vi = e;
        }
        CHECK(ei >= 0 && vi >= 0, "sweep / varied element not found");
        krb_circuit *kc = NULL;
        CHECK(krb_circuit_create(krh_circuit_elements(c), krh_circuit_num_elements(c), n, krh_circuit_is_current(c), &kc) == KRB_OK, "krb_circuit_create");
        krb_plan_opts o;
        krb_default_plan_opts(&o);
        const int32_t varied[1] = {vi};
        o.sweep_element = ei; o.n_varied = 1; o.varied_elements = varied;
        krb_plan *p = NULL;
        CHECK(krb_analyze(kc, KRB_MODE_DC, &o, &p) == KRB_OK && p, "krb_analyze");
        double *vals = (double *)malloc(sizeof(double) * lanes), *x = (double *)calloc((size_t)lanes * num * n, sizeof(double));
        int32_t *status = (int32_t *)calloc(lanes, 4), *iters = (int32_t *)calloc(lanes, 4);
        for (int l = 0; l < lanes; l++) vals[l] = atof(argv[7 + l]);
        krb_overrides ov;
        memset(&ov, 0, sizeof ov);
        ov.n_varied = 1; ov.values = vals;
        CHECK(krb_solve_dc_sweep_sequential(p, &cfg, start, step, num, lanes, &ov, x, iters, status, NULL) == KRB_OK, "krb_solve_dc_sweep_sequential");
        print_rows("lanes", x, (long)lanes * num, n, 0);
        printf(" \"iterations\": [");
        for (int l = 0; l < lanes; l++) printf("%s%d", l ? ", " : "", iters[l]);
        printf("],\n \"status\": [");
        for (int l = 0; l < lanes; l++) printf("%s%d", l ? ", " : "", status[l]);
        printf("]\n");
        free(vals); free(x); free(status); free(iters);
        krb_plan_destroy(p);
        krb_circuit_destroy(kc);
    }
    printf("}\n");
    krh_solver_free(s);
    krh_circuit_free(c);
    return 0;
}
