/*
 * Plain-C client of the drop-in boundary (include/krets_b200.h), compiled by tests/test_c_abi.py with gcc and linked against
 * libkrets_b200.so — what a cgo / Rust-FFI binding sees.
 *
 *   abi_check --layout   prints sizeof / offsetof of every struct of the header as JSON; the test compares it with the layout it
 *                        derives from integration/krets-solver-b200/src/ffi.rs (#[repr(C)] field lists) — the Rust shim cannot be
 *                        compiled in this image (no cargo/rustc), this is its struct-layout check.
 *   abi_check --solve    krb_circuit_create -> krb_analyze -> krb_plan_get_info/pattern -> krb_solve_op / krb_solve_dc_sweep on the
 *                        reference's voltage divider (crates/krets-solver/tests/test_op.rs:78-80: V(out) = 2/3, I(V1) = -1/3000).
 *                        Without a CUDA device the numeric calls must fail with KRB_ERR_NO_DEVICE (no CPU fallback).
 */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "krets_b200.h"

/* the layout the Rust shim (ffi.rs) and INTEGRATION.md document; a change of the header must be a conscious ABI change */
_Static_assert(sizeof(krb_element) == 112 && offsetof(krb_element, value) == 24 && offsetof(krb_element, wave) == 40 &&
               offsetof(krb_element, is) == 96 && offsetof(krb_element, n) == 104, "krb_element layout");
_Static_assert(sizeof(krb_config) == 32 && offsetof(krb_config, maximum_iterations) == 24, "krb_config layout");
_Static_assert(sizeof(krb_plan_opts) == 56 && offsetof(krb_plan_opts, time_step) == 16 && offsetof(krb_plan_opts, varied_elements) == 40 &&
               offsetof(krb_plan_opts, refine_steps) == 52, "krb_plan_opts layout");
_Static_assert(sizeof(krb_io) == 40 && offsetof(krb_io, stream) == 16 && offsetof(krb_io, select) == 24 && offsetof(krb_io, n_select) == 32, "krb_io layout");
_Static_assert(sizeof(krb_overrides) == 16 && offsetof(krb_overrides, values) == 8, "krb_overrides layout");
_Static_assert(sizeof(krb_plan_info) == 72 && offsetof(krb_plan_info, program_ops) == 32 && offsetof(krb_plan_info, executor) == 64, "krb_plan_info layout");

#define F(S, f) printf("%s[\"%s\", %zu]", first++ ? ", " : "", #f, offsetof(S, f))
#define BEGIN(S) printf("%s\"%s\": {\"size\": %zu, \"fields\": [", structs++ ? ",\n " : " ", #S, sizeof(S)); first = 0
#define END() printf("]}")

static int layout(void) {
    int first = 0, structs = 0;
    printf("{\n");
    BEGIN(krb_element); F(krb_element, kind); F(krb_element, ip); F(krb_element, im); F(krb_element, ic); F(krb_element, waveform); F(krb_element, flags);
    F(krb_element, value); F(krb_element, ac_amplitude); F(krb_element, wave); F(krb_element, is); F(krb_element, n); END();
    BEGIN(krb_config); F(krb_config, relative_tolerance); F(krb_config, current_absolute_tolerance); F(krb_config, voltage_absolute_tolerance);
    F(krb_config, maximum_iterations); END();
    BEGIN(krb_plan_opts); F(krb_plan_opts, struct_size); F(krb_plan_opts, exec); F(krb_plan_opts, integrator); F(krb_plan_opts, scratch_slots);
    F(krb_plan_opts, time_step); F(krb_plan_opts, nominal_frequency); F(krb_plan_opts, sweep_element); F(krb_plan_opts, n_varied);
    F(krb_plan_opts, varied_elements); F(krb_plan_opts, threads_per_block); F(krb_plan_opts, refine_steps); END();
    BEGIN(krb_io); F(krb_io, struct_size); F(krb_io, memory); F(krb_io, layout); F(krb_io, reserved); F(krb_io, stream); F(krb_io, select); F(krb_io, n_select); F(krb_io, reserved2); END();
    BEGIN(krb_overrides); F(krb_overrides, n_varied); F(krb_overrides, reserved); F(krb_overrides, values); END();
    printf("\n}\n");
    return 0;
}

#define CHECK(cond, msg)                                                          \
    do {                                                                          \
        if (!(cond)) { fprintf(stderr, "abi_check: %s (%s)\n", msg, krb_last_error_string()); return 1; } \
    } while (0)

static int solve(void) {
    /* V1 in 0 1 ; R1 in out 1000 ; R2 out 0 2000   index map (parser.rs:105-125): I(V1)=0, V(in)=1, V(out)=2 */
    krb_element el[3];
    memset(el, 0, sizeof el);
    el[0].kind = KRB_V; el[0].ip = 1; el[0].im = -1; el[0].ic = 0; el[0].value = 1.0;
    el[1].kind = KRB_R; el[1].ip = 1; el[1].im = 2; el[1].ic = -1; el[1].value = 1000.0;
    el[2].kind = KRB_R; el[2].ip = 2; el[2].im = -1; el[2].ic = -1; el[2].value = 2000.0;
    const uint8_t is_current[3] = {1, 0, 0};
    krb_circuit *c = NULL;
    CHECK(krb_circuit_create(el, 3, 3, is_current, &c) == KRB_OK && c, "krb_circuit_create");
    krb_plan_opts o;
    krb_default_plan_opts(&o);
    CHECK(o.struct_size == (int32_t)sizeof o, "krb_default_plan_opts.struct_size");
    o.sweep_element = 0;
    krb_plan *p = NULL;
    CHECK(krb_analyze(c, KRB_MODE_DC, &o, &p) == KRB_OK && p, "krb_analyze");
    krb_plan_info info;
    CHECK(krb_plan_get_info(p, &info) == KRB_OK && info.n == 3 && info.nnz == 6, "krb_plan_get_info: n = 3, nnz = 6");
    int32_t colptr[4], rowidx[6];
    CHECK(krb_plan_get_pattern(p, colptr, rowidx) == KRB_OK && colptr[3] == 6, "krb_plan_get_pattern");
    krb_config cfg;
    krb_default_config(&cfg);
    CHECK(cfg.maximum_iterations == 300 && cfg.relative_tolerance == 0.001, "krb_default_config (config.rs:26-36)");
    double x[3 * 3] = {0};
    int32_t status[3] = {0};
    const double sweep[3] = {0.0, 1.0, 3.0};
    const int32_t rc = krb_solve_dc_sweep(p, &cfg, sweep, 3, NULL, x, status, NULL);
    if (krb_device_count() <= 0) {
        CHECK(rc == KRB_ERR_NO_DEVICE, "numeric entry points must fail with KRB_ERR_NO_DEVICE without a GPU (no CPU fallback)");
        printf("abi_check: no CUDA device: host-side entry points ok, krb_solve_dc_sweep -> KRB_ERR_NO_DEVICE\n");
    } else {
        CHECK(rc == KRB_OK && !status[0] && !status[1] && !status[2], "krb_solve_dc_sweep");
        for (int l = 0; l < 3; l++) {
            const double v = sweep[l];
            CHECK(fabs(x[3 * l + 1] - v) <= 1e-12 && fabs(x[3 * l + 2] - v * 2.0 / 3.0) <= 1e-12 && fabs(x[3 * l + 0] + v / 3000.0) <= 1e-15,
                  "voltage divider: V(in) = v, V(out) = 2v/3, I(V1) = -v/3000 (test_op.rs:78-80, test_dc.rs:45-53)");
        }
        CHECK(krb_kernel_launch_count() > 0, "krb_kernel_launch_count");
        printf("abi_check: solved 3 sweep points on the GPU through the C ABI: V(out) = %.17g at 1 V\n", x[3 + 2]);
    }
    krb_plan_destroy(p);
    krb_circuit_destroy(c);
    return 0;
}

int main(int argc, char **argv) {
    if (argc > 1 && !strcmp(argv[1], "--layout")) return layout();
    return solve();
}
