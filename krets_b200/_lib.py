"""ctypes loader for libkrets_b200.so (C ABI: include/krets_b200.h + include/krets_b200_host.h).

The library is built in-tree by `make -C krets_b200/csrc` (see __graft_entry__.build()).  There is no Python or CPU
fallback: if the shared library is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KRETS_B200_LIB: profiling knob — load another build of the same library (tools/build_variant.sh), never a fallback
LIB_PATH = os.environ.get("KRETS_B200_LIB") or os.path.join(_HERE, "libkrets_b200.so")


class KrbElement(C.Structure):          # krb_element
    _fields_ = [("kind", C.c_int32), ("ip", C.c_int32), ("im", C.c_int32), ("ic", C.c_int32), ("waveform", C.c_int32),
                ("flags", C.c_int32), ("value", C.c_double), ("ac_amplitude", C.c_double), ("wave", C.c_double * 7),
                ("is_", C.c_double), ("n", C.c_double)]


class KrbConfig(C.Structure):           # krb_config
    _fields_ = [("relative_tolerance", C.c_double), ("current_absolute_tolerance", C.c_double),
                ("voltage_absolute_tolerance", C.c_double), ("maximum_iterations", C.c_int32)]


class KrbPlanOpts(C.Structure):         # krb_plan_opts
    _fields_ = [("struct_size", C.c_int32), ("exec", C.c_int32), ("integrator", C.c_int32), ("scratch_slots", C.c_int32),
                ("time_step", C.c_double), ("nominal_frequency", C.c_double), ("sweep_element", C.c_int32),
                ("n_varied", C.c_int32), ("varied_elements", C.POINTER(C.c_int32)), ("threads_per_block", C.c_int32),
                ("refine_steps", C.c_int32)]


class KrbPlanInfo(C.Structure):         # krb_plan_info
    _fields_ = [("n", C.c_int32), ("nnz", C.c_int32), ("nnz_lu", C.c_int32), ("mode", C.c_int32), ("nonlinear", C.c_int32),
                ("scratch_slots", C.c_int32), ("spill_planes", C.c_int32), ("n_inputs", C.c_int32), ("program_ops", C.c_int64),
                ("flops_per_point", C.c_int64), ("hbm_bytes_per_point", C.c_int64), ("alg_bytes_per_point", C.c_int64),
                ("executor", C.c_int32), ("reserved", C.c_int32)]


class KrbIO(C.Structure):               # krb_io
    _fields_ = [("struct_size", C.c_int32), ("memory", C.c_int32), ("layout", C.c_int32), ("reserved", C.c_int32),
                ("stream", C.c_void_p), ("select", C.POINTER(C.c_int32)), ("n_select", C.c_int32), ("reserved2", C.c_int32)]


class KrbOverrides(C.Structure):        # krb_overrides
    _fields_ = [("n_varied", C.c_int32), ("reserved", C.c_int32), ("values", C.c_void_p)]


# every symbol the two headers declare: name -> (restype, argtypes)
_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pi32, _pf64 = C.POINTER(C.c_int32), C.POINTER(C.c_double)
SYMBOLS = {
    # include/krets_b200.h
    "krb_default_config": (None, [C.POINTER(KrbConfig)]),
    "krb_default_plan_opts": (None, [C.POINTER(KrbPlanOpts)]),
    "krb_default_io": (None, [C.POINTER(KrbIO)]),
    "krb_circuit_create": (_i32, [C.POINTER(KrbElement), _i32, _i32, C.POINTER(C.c_uint8), C.POINTER(_vp)]),
    "krb_circuit_destroy": (None, [_vp]),
    "krb_analyze": (_i32, [_vp, _i32, C.POINTER(KrbPlanOpts), C.POINTER(_vp)]),
    "krb_plan_destroy": (None, [_vp]),
    "krb_plan_get_info": (_i32, [_vp, C.POINTER(KrbPlanInfo)]),
    "krb_plan_get_pattern": (_i32, [_vp, _pi32, _pi32]),
    "krb_plan_get_pivot_order": (_i32, [_vp, _pi32, _pi32]),
    "krb_plan_program_size": (_i64, [_vp, _i32]),
    "krb_plan_get_program": (_i32, [_vp, _i32, C.POINTER(C.c_uint32)]),
    "krb_plan_jit_source": (_i64, [_vp, _i32, C.c_char_p, _i64]),
    "krb_plan_jit_compile": (_i32, [_vp, C.POINTER(_i64)]),
    "krb_solve_op": (_i32, [_vp, C.POINTER(KrbConfig), _i64, C.POINTER(KrbOverrides), _vp, _vp, _vp, C.POINTER(KrbIO)]),
    "krb_solve_dc_sweep": (_i32, [_vp, C.POINTER(KrbConfig), _vp, _i64, C.POINTER(KrbOverrides), _vp, _vp, C.POINTER(KrbIO)]),
    "krb_solve_dc_sweep_sequential": (_i32, [_vp, C.POINTER(KrbConfig), _f64, _f64, _i64, _i64, C.POINTER(KrbOverrides), _vp, _vp, _vp, C.POINTER(KrbIO)]),
    "krb_solve_ac": (_i32, [_vp, C.POINTER(KrbConfig), _vp, _i64, C.POINTER(KrbOverrides), _vp, _i64, _vp, _vp, C.POINTER(KrbIO)]),
    "krb_solve_transient": (_i32, [_vp, _vp, C.POINTER(KrbConfig), _i64, _i64, C.POINTER(KrbOverrides), _vp, _vp, C.POINTER(KrbIO)]),
    "krb_last_error_string": (C.c_char_p, []),
    "krb_version": (C.c_char_p, []),
    "krb_device_count": (_i32, []),
    "krb_kernel_launch_count": (_i64, []),
    "krb_last_kernel_ms": (_f64, []),
    # include/krets_b200_host.h
    "krh_parse_circuit_description": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "krh_parse_circuit_description_file": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "krh_circuit_free": (None, [_vp]),
    "krh_circuit_num_unknowns": (_i32, [_vp]),
    "krh_circuit_unknown_name": (C.c_char_p, [_vp, _i32]),
    "krh_circuit_index_of": (_i32, [_vp, C.c_char_p]),
    "krh_circuit_num_elements": (_i32, [_vp]),
    "krh_circuit_element_identifier": (C.c_char_p, [_vp, _i32]),
    "krh_circuit_elements": (C.POINTER(KrbElement), [_vp]),
    "krh_circuit_is_current": (C.POINTER(C.c_uint8), [_vp]),
    "krh_circuit_set_device_models": (_i32, [_vp, _i32]),
    "krh_generate_frequencies": (_i64, [_i32, C.c_uint32, _f64, _f64, _pf64, _i64]),
    "krh_dc_num_steps": (_i32, [_f64, _f64, _f64]),
    "krh_transient_num_steps": (_i32, [_f64, _f64]),
    "krh_solver_new": (_i32, [_vp, C.POINTER(KrbConfig), C.POINTER(_vp)]),
    "krh_solver_free": (None, [_vp]),
    "krh_solver_solve_op": (_i32, [_vp, C.POINTER(_vp)]),
    "krh_solver_solve_dc": (_i32, [_vp, C.c_char_p, _f64, _f64, _f64, C.POINTER(_vp)]),
    "krh_solver_solve_dc_with": (_i32, [_vp, C.c_char_p, _f64, _f64, _f64, _i32, C.POINTER(_vp)]),
    "krh_solver_solve_ac": (_i32, [_vp, _i32, C.c_uint32, _f64, _f64, C.POINTER(_vp)]),
    "krh_solver_solve_transient": (_i32, [_vp, _f64, _f64, C.POINTER(_vp)]),
    "krh_solver_solve_transient_with": (_i32, [_vp, _f64, _f64, _i32, C.POINTER(_vp)]),
    "krh_result_kind": (_i32, [_vp]),
    "krh_result_rows": (_i64, [_vp]),
    "krh_result_num_unknowns": (_i32, [_vp]),
    "krh_result_data": (_pf64, [_vp]),
    "krh_result_axis": (_pf64, [_vp]),
    "krh_result_free": (None, [_vp]),
    "krh_last_error_string": (C.c_char_p, []),
}
# introspection helpers used by the CPU-side program emulator in tests/ (not in the public headers)
_DEBUG_SYMBOLS = {
    "krb_plan_pool_size": (_i64, [_vp]),
    "krb_plan_eval_pool": (_i32, [_vp, _pf64]),
    "krb_plan_wave_count": (_i32, [_vp]),
    "krb_plan_wave_value": (_i32, [_vp, _i32, _f64, _pf64]),
    "krb_plan_input_layout": (_i32, [_vp, _pi32, _pi32, _pi32]),
    "krb_plan_has_step_sections": (_i32, [_vp]),
    "krb_plan_state_columns": (_i32, [_vp]),
    "krb_plan_integrator": (_i32, [_vp]),
    "krb_plan_jit_active": (_i32, [_vp]),
    "krb_plan_repivoted": (_i32, [_vp]),
    "krb_plan_chain_size": (_i64, [_vp]),
    "krb_plan_get_chain": (_i32, [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), _pi32]),
    "krb_plan_front_array": (_i64, [_vp, _i32, _pi32]),
    "krb_plan_front_info": (_i32, [_vp, C.POINTER(C.c_int64)]),
}

_lib = None


def lib():
    """Load libkrets_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C krets_b200/csrc` "
                              "(krets_b200 has no Python/CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in {**SYMBOLS, **_DEBUG_SYMBOLS}.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
