//! Raw bindings of include/krets_b200.h (field order and widths must match the header exactly).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_void};

pub const KRB_OK: i32 = 0;
pub const KRB_ERR_ELEMENT_NOT_FOUND: i32 = 3;
pub const KRB_ERR_INVALID_ELEMENT_FORMAT: i32 = 4;
pub const KRB_ERR_MAX_ITERATIONS: i32 = 5;
pub const KRB_ERR_DECOMPOSITION_FAILED: i32 = 6;
pub const KRB_ERR_UNSUPPORTED: i32 = 7;

pub const KRB_LANE_ZERO_PIVOT: i32 = 1;
pub const KRB_LANE_NOT_CONVERGED: i32 = 2;
pub const KRB_LANE_NONFINITE: i32 = 4;
pub const KRB_LANE_INACCURATE: i32 = 8;

pub const KRB_R: i32 = 0;
pub const KRB_C: i32 = 1;
pub const KRB_L: i32 = 2;
pub const KRB_V: i32 = 3;
pub const KRB_I: i32 = 4;
pub const KRB_D: i32 = 5;
pub const KRB_NMOS: i32 = 6;
pub const KRB_WAVE_NONE: i32 = 0;
pub const KRB_WAVE_PULSE: i32 = 1;
pub const KRB_WAVE_SIN: i32 = 2;
pub const KRB_ELEM_FLAG_G2: i32 = 1;
pub const KRB_ELEM_FLAG_LEVEL1: i32 = 2; // extension: NMOS stamps the level-1 companion model (krets_b200.h)

pub const KRB_MODE_DC: i32 = 0;
pub const KRB_MODE_AC: i32 = 1;
pub const KRB_MODE_TRAN: i32 = 2;
pub const KRB_INTEGRATOR_BE_KRETS: i32 = 0;
pub const KRB_EXEC_FUSED: i32 = 0;
pub const KRB_MEM_HOST: i32 = 0;
pub const KRB_LAYOUT_POINT_MAJOR: i32 = 0;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct krb_element {
    pub kind: i32,
    pub ip: i32,
    pub im: i32,
    pub ic: i32,
    pub waveform: i32,
    pub flags: i32,
    pub value: f64,
    pub ac_amplitude: f64,
    pub wave: [f64; 7],
    pub is: f64,
    pub n: f64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct krb_config {
    pub relative_tolerance: f64,
    pub current_absolute_tolerance: f64,
    pub voltage_absolute_tolerance: f64,
    pub maximum_iterations: i32,
}

#[repr(C)]
pub struct krb_plan_opts {
    pub struct_size: i32,
    pub exec: i32,
    pub integrator: i32,
    pub scratch_slots: i32,
    pub time_step: f64,
    pub nominal_frequency: f64,
    pub sweep_element: i32,
    pub n_varied: i32,
    pub varied_elements: *const i32,
    pub threads_per_block: i32,
    pub refine_steps: i32,
}

#[repr(C)]
pub struct krb_io {
    pub struct_size: i32,
    pub memory: i32,
    pub layout: i32,
    pub reserved: i32,
    pub stream: *mut c_void,
    pub select: *const i32, // extension: deliver only these unknowns (krets_b200.h)
    pub n_select: i32,
    pub reserved2: i32,
}

#[repr(C)]
pub struct krb_overrides {
    pub n_varied: i32,
    pub reserved: i32,
    pub values: *const f64,
}

#[repr(C)]
pub struct krb_circuit {
    _private: [u8; 0],
}
#[repr(C)]
pub struct krb_plan {
    _private: [u8; 0],
}

unsafe extern "C" {
    pub fn krb_default_plan_opts(o: *mut krb_plan_opts);
    pub fn krb_default_io(io: *mut krb_io);
    pub fn krb_circuit_create(
        elements: *const krb_element,
        n_elements: i32,
        n_unknowns: i32,
        unknown_is_current: *const u8,
        out: *mut *mut krb_circuit,
    ) -> i32;
    pub fn krb_circuit_destroy(c: *mut krb_circuit);
    pub fn krb_analyze(c: *const krb_circuit, mode: i32, opts: *const krb_plan_opts, out: *mut *mut krb_plan) -> i32;
    pub fn krb_plan_destroy(p: *mut krb_plan);
    pub fn krb_solve_op(
        plan_dc: *mut krb_plan,
        cfg: *const krb_config,
        lanes: i64,
        ov: *const krb_overrides,
        x_out: *mut f64,
        iterations: *mut i32,
        lane_status: *mut i32,
        io: *const krb_io,
    ) -> i32;
    pub fn krb_solve_dc_sweep(
        plan_dc: *mut krb_plan,
        cfg: *const krb_config,
        sweep_values: *const f64,
        lanes: i64,
        ov: *const krb_overrides,
        x_out: *mut f64,
        lane_status: *mut i32,
        io: *const krb_io,
    ) -> i32;
    pub fn krb_solve_dc_sweep_sequential(
        plan_dc: *mut krb_plan,
        cfg: *const krb_config,
        start: f64,
        step_size: f64,
        n_points: i64,
        lanes: i64,
        ov: *const krb_overrides,
        x_out: *mut f64,
        iterations: *mut i32,
        lane_status: *mut i32,
        io: *const krb_io,
    ) -> i32;
    pub fn krb_solve_ac(
        plan_ac: *mut krb_plan,
        cfg: *const krb_config,
        frequencies: *const f64,
        lanes: i64,
        ov: *const krb_overrides,
        op_solution: *const f64,
        lanes_op: i64,
        x_out: *mut f64,
        lane_status: *mut i32,
        io: *const krb_io,
    ) -> i32;
    pub fn krb_solve_transient(
        plan_dc: *mut krb_plan,
        plan_tran: *mut krb_plan,
        cfg: *const krb_config,
        n_steps: i64,
        lanes: i64,
        ov: *const krb_overrides,
        x_out: *mut f64,
        lane_status: *mut i32,
        io: *const krb_io,
    ) -> i32;
    pub fn krb_last_error_string() -> *const c_char;
}
