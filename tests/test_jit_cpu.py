"""CPU tier: circuit-specialised lane kernels (krets_b200/csrc/vm_jit.cpp).

The text the library hands to NVRTC on the GPU box is (a) compiled here by NVRTC for sm_100a (no device needed) and (b) built as
host C++ behind tests/jit_host_shim.h and run against the oracle and the numpy emulator of the interpreter: one statement per
lane-program operation, same order, same arithmetic — so (b) holds the generated statements to the reference's results."""
import ctypes as C

import numpy as np
import pytest

import netgen
from helpers import assert_close
from jit_host import JitHost, jit_source
from test_host_cpu import kcirc, ocirc
from vm_emulator import Emulator, input_layout

import krets_b200 as kb
from krets_b200 import _lib, api


def test_jit_source_one_statement_per_operation(golden):
    kc = kcirc(golden, "rectifier")
    plan = kb.Plan(kc, api.MODE_TRAN, time_step=50e-6)
    src = jit_source(plan)
    assert src and "krb_jit_vm" in src and "KRB_JIT_HOST 1" not in src
    body = src[src.index("if (prog == 0) {"):]
    n_stmt = sum(1 for ln in body.splitlines() if ln.startswith(("s[", "c.G[", "k_xout(")))
    n_ops = 0
    for sec in range(3):
        prog = plan.program(sec)
        i = 0
        while i < len(prog):
            opc = int(prog[i][0]) & 0xFF
            if opc == 0:
                break
            if opc == 21:      # LDK: constant in the next slot
                i += 1
            if opc not in (22, 23):   # NOP, WAIT have no statement
                n_ops += 1
            i += 1
    assert n_stmt == n_ops > 50
    # a ladder on the chain executor and a program beyond KRB_VM_JIT_MAX_OPS are not specialised
    ladder = kb.parse_circuit_description(netgen.resistor_ladder(300))
    assert jit_source(kb.Plan(ladder, api.MODE_DC, sweep_element=0)) is None
    assert jit_source(kb.Plan(kb.parse_circuit_description(netgen.resistor_ladder(5000)), api.MODE_DC, sweep_element=0, force_lane_vm=True)) is None


@pytest.mark.parametrize("key,mode", [("rectifier", api.MODE_TRAN), ("diode_iv_curve", api.MODE_DC), ("low_pass_filter", api.MODE_AC),
                                      ("dual_rc_ladder", api.MODE_TRAN), ("farid_n_najm", api.MODE_DC)])
def test_jit_source_compiles_with_nvrtc(golden, key, mode):
    """NVRTC accepts the generated text for sm_100a (the step the GPU box performs at first use)."""
    kw = dict(time_step=50e-6) if mode == api.MODE_TRAN else {}
    if mode == api.MODE_AC:
        kw = dict(nominal_frequency=1e3, force_lane_vm=True)
    plan = kb.Plan(kcirc(golden, key), mode, **kw)
    nbytes = C.c_int64()
    rc = _lib.lib().krb_plan_jit_compile(plan._h, C.byref(nbytes))
    assert rc == 0, kb.last_error() if hasattr(kb, "last_error") else _lib.lib().krb_last_error_string()
    assert nbytes.value > 1000


@pytest.mark.parametrize("slots", [8, 48])
def test_jit_host_op_small_circuits(oracle, golden, slots):
    for key in ["farid_n_najm", "test_case_1", "voltage_divider", "basic_001", "high_pass_filter", "low_pass_filter", "inverter"]:
        kc, oc = kcirc(golden, key), ocirc(oracle, golden, key)
        plan = kb.Plan(kc, api.MODE_DC, scratch_slots=slots, force_lane_vm=True)
        X = JitHost(plan, 1).run(np.zeros((0, 1)))
        assert_close(X[0][:, 0], oc.solve_op(), what=key)


def test_jit_host_nr_diode(oracle, golden):
    kc, oc = kcirc(golden, "diode_iv_curve"), ocirc(oracle, golden, "diode_iv_curve")
    plan = kb.Plan(kc, api.MODE_DC)
    jh = JitHost(plan, 1)
    X = jh.run(np.zeros((0, 1)))
    ref, it = oc.solve_op(return_iterations=True)
    assert_close(X[0][:, 0], ref, what="diode OP")
    assert jh.iters[0] == it and jh.status[0] == 0


def test_jit_host_dc_sweep_sequential_warm_start(oracle, golden):
    """The specialised kernel on the sequential DC-sweep schedule (dc.rs:58,112 warm start) == the reference's iteration history."""
    from vm_emulator import Emulator, run_sequential_sweep
    kc, oc = kcirc(golden, "diode_iv_curve"), ocirc(oracle, golden, "diode_iv_curve")
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("V1"))
    X, jh, its = run_sequential_sweep(JitHost, plan, 0.0, 0.125, 17)
    assert_close(X[:, : plan.n, 0], oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=True), rtol=1e-9, what="sequential sweep (specialised kernel)")
    Xe, em, ite = run_sequential_sweep(Emulator, plan, 0.0, 0.125, 17)
    assert_close(X[:, : plan.n, 0], Xe[:, : plan.n, 0], rtol=1e-12, what="specialised kernel vs interpreter emulator")
    assert np.array_equal(its, ite) and not jh.status.any()


def test_jit_host_dc_sweep_with_spills(oracle):
    """A 100-section ladder forced onto the lane VM with 12 scratch cells: LDG/STG planes, LDI sweep input."""
    text = netgen.resistor_ladder(100)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=0, scratch_slots=12, force_lane_vm=True)
    vals = np.linspace(-2.0, 3.0, 7)
    jh = JitHost(plan, len(vals))
    X = jh.run(vals[None, :])
    ref = oc.solve_dc("V1", -2.0, 3.0, 5.0 / 6.0)
    assert_close(X[0].T, ref, what="ladder DC sweep")
    assert not jh.status.any()
    em = Emulator(plan, len(vals))
    Xe = em.run(vals[None, :])
    assert_close(X[0].T, Xe[0].T, rtol=1e-12, what="specialised kernel vs emulated interpreter")   # the emulator does not fuse d - a*b


def test_jit_host_ac(oracle, golden):
    for key, n in [("low_pass_filter", 0), ("high_pass_filter", 0), ("rc", 40)]:
        text = netgen.rc_ladder(n) if key == "rc" else golden["netlists"][key]
        kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
        freqs = oracle.generate_frequencies("Decade", 3, 1.0, 1e7)
        plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e3, scratch_slots=14, force_lane_vm=True)
        X = JitHost(plan, len(freqs)).run(freqs[None, :])
        _, ref = oc.solve_ac(freqs)
        assert_close(X[0].T, ref, what=key)


def test_jit_host_transient_linear_and_nonlinear(oracle, golden):
    for key, h, stop, rtol in [("dual_rc_ladder", 50e-6, 5e-3, 1e-9), ("high_pass_filter_transient", 10e-6, 1e-3, 1e-9), ("rectifier", 50e-6, 5e-3, 1e-7)]:
        kc, oc = kcirc(golden, key), ocirc(oracle, golden, key)
        pdc = kb.Plan(kc, api.MODE_DC, force_lane_vm=True)
        ptr = kb.Plan(kc, api.MODE_TRAN, time_step=h, force_lane_vm=True)
        t, ref = oc.solve_tran(h, stop)
        rows = len(t)
        X = JitHost(pdc, 1, rows=rows).run(np.zeros((0, 1)), row_begin=0, row_end=1)
        jh = JitHost(ptr, 1, rows=rows, h=h)
        X = jh.run(np.zeros((0, 1)), X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
        assert_close(X[:, :, 0], ref, rtol=rtol, what=key)
        assert not jh.status.any()


def test_jit_host_monte_carlo_transient(oracle, golden):
    text = golden["netlists"]["dual_rc_ladder"]
    kc = kb.parse_circuit_description(text)
    names = ["R1", "R2", "C1", "C2"]
    varied = [kc.element_index(nm) for nm in names]
    lanes, h, nsteps = 4, 5e-6, 30
    vals = np.array([[kc.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(lanes)] for j, e in enumerate(varied)])
    pdc = kb.Plan(kc, api.MODE_DC, varied=varied, force_lane_vm=True)
    ptr = kb.Plan(kc, api.MODE_TRAN, varied=varied, time_step=h, force_lane_vm=True)
    _, _, iov = input_layout(ptr)
    IN = np.zeros((ptr.info.n_inputs, lanes))
    for k, pl in enumerate(iov):
        IN[pl] = vals[k]
    rows = nsteps + 1
    X = JitHost(pdc, lanes, rows=rows).run(IN, row_begin=0, row_end=1)
    X = JitHost(ptr, lanes, rows=rows, h=h).run(IN, X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
    for l in range(lanes):
        oc = oracle.Circuit(text)
        for j, nm in enumerate(names):
            oc.set_element_value(oc.element_index(nm), vals[j, l])
        _, ref = oc.solve_tran(h, nsteps * h)
        assert_close(X[:, :, l], ref, what=f"lane {l}")


def test_jit_host_chunked_complex_mesh(oracle):
    """A program beyond the inline limit (1200 operations per section) is cut into __noinline__ chunks that hand live cells over
    through the per-thread array: a 5 x 6 RLC mesh in AC (complex arithmetic, fill-in, spill planes with 12 scratch cells)."""
    text = netgen.rlc_mesh(5, 6, transient=False)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = np.array([1.0, 1e3, 1e6, 1e9])
    plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e4, scratch_slots=12, force_lane_vm=True)
    src = jit_source(plan)
    assert plan.info.program_ops > 1200 and src.count("__noinline__ void krb_s0_c") >= 4
    X = JitHost(plan, len(freqs)).run(freqs[None, :])
    _, ref = oc.solve_ac(freqs)
    assert_close(X[0].T, ref, what="mesh AC, chunked specialised kernel")
    nbytes = C.c_int64()
    assert _lib.lib().krb_plan_jit_compile(plan._h, C.byref(nbytes)) == 0 and nbytes.value > 1000


def test_vmparams_text_matches_the_interpreter_struct():
    """The specialised kernels take krb::VmParams by value through a struct re-declared in vm_jit_prelude.inc (NVRTC sees no project
    headers; the generated text static_asserts the size).  Field order and kinds must match vm_kernel.cuh declaration by declaration."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "krets_b200", "csrc")

    def fields(path):
        text = open(os.path.join(root, path)).read()
        body = re.search(r"struct VmParams \{(.*?)\n\};", text, re.S).group(1)
        out = []
        for decl in re.sub(r"//[^\n]*", "", body).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            base = "ptr" if "*" in decl.split(",")[0] else ("ll" if "long long" in decl else "double" if decl.startswith("double") else "int")
            for nm in decl.split(","):
                kind = "ptr" if "*" in nm or (base == "ptr" and nm is decl.split(",")[0]) else base
                out.append((re.findall(r"[A-Za-z_][A-Za-z_0-9]*", nm)[-1], kind))
        return out

    a, b = fields("vm_kernel.cuh"), fields("vm_jit_prelude.inc")
    assert a == b and len(a) == 25, (a, b)
