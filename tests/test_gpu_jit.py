"""GPU tier: circuit-specialised lane kernels (krets_b200/csrc/vm_jit.cpp, NVRTC at first use) against the oracle and against the
interpreter (KRB_VM_JIT=0) on the same inputs — the two execute the same operations in the same order and must agree bit for bit."""
import os

import numpy as np
import pytest

import netgen
from helpers import assert_close
from test_host_cpu import kcirc, ocirc

import krets_b200 as kb
from krets_b200 import _lib, api

pytestmark = pytest.mark.gpu


def _with_jit(flag, fn):
    old = os.environ.get("KRB_VM_JIT")
    os.environ["KRB_VM_JIT"] = "1" if flag else "0"
    try:
        return fn()
    finally:
        if old is None:
            del os.environ["KRB_VM_JIT"]
        else:
            os.environ["KRB_VM_JIT"] = old


def _active(plan):
    return int(_lib.lib().krb_plan_jit_active(plan._h))


def test_rectifier_monte_carlo_specialised_vs_interpreter(oracle, golden):
    text = golden["netlists"]["rectifier"]
    circuit = kb.parse_circuit_description(text)
    names = ["C1", "R1"]
    varied = [circuit.element_index(nm) for nm in names]
    lanes, nsteps, h = 200, 300, 50e-6
    vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 2) for l in range(lanes)] for j, e in enumerate(varied)])

    def run():
        pdc = kb.Plan(circuit, api.MODE_DC, varied=varied)
        ptr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h)
        out, status = ptr.solve_transient(pdc, nsteps, lanes, varied_values=vals)[:2]
        return out, status, _active(pdc), _active(ptr)

    out_j, st_j, a0, a1 = _with_jit(True, run)
    out_i, st_i, b0, b1 = _with_jit(False, run)
    assert (a0, a1) == (1, 1) and (b0, b1) == (0, 0)
    assert not st_j.any() and not st_i.any()
    assert np.array_equal(out_j, out_i)
    for l in (0, 77, 199):
        oc = oracle.Circuit(text)
        for j, nm in enumerate(names):
            oc.set_element_value(oc.element_index(nm), vals[j, l])
        _, ref = oc.solve_tran(h, nsteps * h)
        assert_close(out_j[l], ref, rtol=1e-6, what=f"rectifier lane {l} (NR per step; exp() differs in the last ulp)")


def test_diode_op_iterations_specialised(oracle, golden):
    kc, oc = kcirc(golden, "diode_iv_curve"), ocirc(oracle, golden, "diode_iv_curve")

    def run():
        plan = kb.Plan(kc, api.MODE_DC)
        out, status, iters = plan.solve_op(return_iterations=True)
        return out, status, iters, _active(plan)

    out_j, st_j, it_j, a = _with_jit(True, run)
    out_i, st_i, it_i, b = _with_jit(False, run)
    ref, it = oc.solve_op(return_iterations=True)
    assert (a, b) == (1, 0)
    assert_close(out_j[0], ref, what="diode OP")
    assert it_j[0] == it_i[0] == it and st_j[0] == st_i[0] == 0
    assert np.array_equal(out_j, out_i)
    # NR that cannot converge in two iterations reports it on both paths
    cfg = kb.SolverConfig(maximum_iterations=2)
    for flag in (True, False):
        st = _with_jit(flag, lambda: kb.Plan(kc, api.MODE_DC).solve_op(cfg=cfg)[1])
        assert st[0] & 2


def test_ac_complex_and_dc_sweep_with_spills_specialised(oracle, golden):
    text = netgen.rc_ladder(40)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = oracle.generate_frequencies("Decade", 25, 1.0, 1e7)

    def run_ac():
        plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e3, scratch_slots=14, force_lane_vm=True)
        out, status = plan.solve_ac(freqs)[:2]
        return out, status, _active(plan)

    out_j, st_j, a = _with_jit(True, run_ac)
    out_i, st_i, b = _with_jit(False, run_ac)
    assert (a, b) == (1, 0) and not st_j.any()
    _, ref = oc.solve_ac(freqs)
    assert_close(out_j, ref, what="rc ladder AC, specialised complex kernel")
    assert np.array_equal(out_j, out_i)

    text = netgen.resistor_ladder(100)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    vals = np.linspace(-2.0, 3.0, 333)

    def run_dc():
        plan = kb.Plan(kc, api.MODE_DC, sweep_element=0, scratch_slots=12, force_lane_vm=True)
        out, status = plan.solve_dc_sweep(vals)
        return out, status, _active(plan), plan.info.spill_planes

    out_j, st_j, a, planes = _with_jit(True, run_dc)
    out_i, st_i, b, _ = _with_jit(False, run_dc)
    assert (a, b) == (1, 0) and planes > 0 and not st_j.any()
    ref = oc.solve_dc("V1", -2.0, 3.0, 5.0 / 332.0)
    assert ref.shape[0] == len(vals)
    assert_close(out_j, ref, what="ladder DC sweep through spill planes, specialised kernel")
    assert np.array_equal(out_j, out_i)


def test_long_programs_stay_on_the_interpreter():
    kc = kb.parse_circuit_description(netgen.resistor_ladder(3000))
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=0, force_lane_vm=True)
    assert plan.info.program_ops > 2500
    out, status = plan.solve_dc_sweep(np.linspace(0.0, 1.0, 64))
    assert _active(plan) == 0 and not status.any()


def test_mid_size_program_chunked_kernel(oracle):
    """A 7 x 7 RLC mesh forced onto the lane VM (4.8 k operations: beyond the always-specialise limit, cut into __noinline__ chunks) is
    specialised once its first solve has >= 32768 points; a small batch of the same plan stays on the interpreter.  Both agree bit for bit."""
    text = netgen.rlc_mesh(7, 7, transient=True)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    names = ["Rs", "RL", "Rh3"]
    varied = [kc.element_index(nm) for nm in names]
    lanes = 32768
    vals = np.array([[kc.element_value(e) * netgen.mc_factor(l, j, 3) for l in range(lanes)] for j, e in enumerate(varied)])

    def run(n):
        plan = kb.Plan(kc, api.MODE_DC, varied=varied, force_lane_vm=True)
        assert 2500 < plan.info.program_ops <= 20000
        v = vals[:, :n].copy()
        plan.solve_op(n, varied_values=v)                    # first solve: upload (+ NVRTC)
        out, status = plan.solve_op(n, varied_values=v)
        return out, status, _active(plan), kb.last_kernel_ms()

    out_j, st_j, a, ms_j = _with_jit(True, lambda: run(lanes))
    out_i, st_i, b, ms_i = _with_jit(False, lambda: run(lanes))
    out_s, st_s, c, _ = _with_jit(True, lambda: run(64))
    assert (a, b, c) == (1, 0, 0)
    assert not st_j.any() and not st_i.any()
    assert np.array_equal(out_j, out_i) and np.array_equal(out_s, out_i[:64])
    for l in (0, 12345, lanes - 1):
        o = oracle.Circuit(text)
        for j, nm in enumerate(names):
            o.set_element_value(o.element_index(nm), vals[j, l])
        assert_close(out_j[l], o.solve_op(), what=f"mesh OP lane {l}")
    print(f"\n7x7 mesh OP x {lanes} lanes: specialised (chunked) {ms_j:.3f} ms, interpreter {ms_i:.3f} ms (device time of the second solve)")


def test_cubin_cache_directory(golden, tmp_path, monkeypatch):
    """KRB_VM_JIT_CACHE: the compiled kernel is published there once; a second plan of the same program loads it instead of calling NVRTC
    (same results), and a corrupt file falls back to NVRTC."""
    import time
    monkeypatch.setenv("KRB_VM_JIT_CACHE", str(tmp_path))
    monkeypatch.setenv("KRB_VM_JIT", "1")
    kc = kcirc(golden, "diode_iv_curve")

    def run():
        t0 = time.time()
        plan = kb.Plan(kc, api.MODE_DC)
        out, status = plan.solve_op()
        return out, _active(plan), time.time() - t0

    out1, a1, t1 = run()
    files = sorted(tmp_path.glob("*.cubin"))
    assert a1 == 1 and len(files) == 1 and files[0].stat().st_size > 1000
    out2, a2, t2 = run()
    assert a2 == 1 and np.array_equal(out1, out2) and len(list(tmp_path.glob("*"))) == 1
    files[0].write_bytes(b"not a cubin" * 20)
    out3, a3, _ = run()
    assert a3 == 1 and np.array_equal(out1, out3) and files[0].stat().st_size > 1000      # recompiled and republished
    print(f"\nfirst solve {t1 * 1e3:.0f} ms (NVRTC), from the cache {t2 * 1e3:.0f} ms")
