"""GPU tier: level-1 NMOS companion stamps (EXTENSION) through the public API against the oracle's extension
(pinned to an independent KCL solve by tests/test_nmos_level1.py), on the specialised kernel and on the interpreter."""
import os

import numpy as np
import pytest

import netgen
from helpers import assert_close
from test_nmos_level1 import AMP, AMP_TRAN, circuits

import krets_b200 as kb
from krets_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture()
def level1(oracle):
    oracle.set_device_models(1)
    yield
    oracle.set_device_models(0)


@pytest.fixture(params=["specialised", "interpreter"])
def executor(request):
    old = os.environ.get("KRB_VM_JIT")
    os.environ["KRB_VM_JIT"] = "1" if request.param == "specialised" else "0"
    yield request.param
    if old is None:
        del os.environ["KRB_VM_JIT"]
    else:
        os.environ["KRB_VM_JIT"] = old


def test_solver_op_dc_transient_ac(oracle, level1, executor):
    kc, oc = circuits(oracle, AMP)
    solver = kb.Solver(kc)
    op = solver.solve(kb.Analysis.Op()).into_op()
    ref = oc.solve_op()
    assert_close(np.array([op[nm] for nm in oc.names]), ref, what="common-source OP")
    dc = solver.solve(kb.Analysis.Dc(kb.DcAnalysis("VG", 0.0, 2.25, 0.25)))
    assert_close(dc.array, oc.solve_dc("VG", 0.0, 2.25, 0.25, warm_start=True), rtol=1e-8, what="gate sweep vs the sequential warm-started oracle (dc.rs:58)")
    dc = solver.solve(kb.Analysis.Dc(kb.DcAnalysis("VG", 0.0, 2.25, 0.25, mode=api.DC_PARALLEL_COLD_START)))
    cold = oc.solve_dc("VG", 0.0, 2.25, 0.25, warm_start=False)
    assert_close(dc.array, cold, rtol=1e-8, what="gate sweep, parallel lanes vs cold-start oracle")
    ac = solver.solve(kb.Analysis.Ac(kb.AcAnalysis(kb.AcSweep.Decade(5), 10.0, 1e6)))
    fr, refa = oc.solve_ac(oracle.generate_frequencies("Decade", 5, 10.0, 1e6))
    assert ac.axis.tobytes() == fr.tobytes()
    assert_close(ac.array, refa, what="common-source AC (small-signal g_m, g_ds at the operating point)")
    d = oc.names.index("V(d)")
    assert abs(refa[0, d]) > 2.0          # the stage has gain
    kc, oc = circuits(oracle, AMP_TRAN)
    tr = kb.Solver(kc).solve(kb.Analysis.Transient(kb.TransientAnalysis(10e-6, 2e-3)))
    t, reft = oc.solve_tran(10e-6, 2e-3)
    assert tr.array.shape == reft.shape
    assert_close(tr.array, reft, what="common-source transient")
    # flag off: the reference's no-op on the same netlist
    oracle.set_device_models(0)
    kc0 = kb.parse_circuit_description(AMP)
    op0 = kb.Solver(kc0).solve(kb.Analysis.Op()).into_op()
    assert_close(np.array([op0[nm] for nm in oc.names if nm in op0]), oracle.Circuit(AMP).solve_op(), what="NMOS no-op (reference)")


def test_monte_carlo_lanes_and_nonconvergence(oracle, level1, executor):
    kc, oc = circuits(oracle, AMP)
    names = ["RD", "RS"]
    varied = [kc.element_index(nm) for nm in names]
    lanes = 96
    vals = np.array([[kc.element_value(e) * netgen.mc_factor(l, j, 2) for l in range(lanes)] for j, e in enumerate(varied)])
    plan = kb.Plan(kc, api.MODE_DC, varied=varied)
    out, status, iters = plan.solve_op(lanes, varied_values=vals, return_iterations=True)
    assert not status.any()
    for l in (0, 31, 95):
        o = oracle.Circuit(AMP)
        for j, nm in enumerate(names):
            o.set_element_value(o.element_index(nm), vals[j, l])
        ref, it = o.solve_op(return_iterations=True)
        assert_close(out[l], ref, what=f"lane {l}")
        assert iters[l] == it
    # a lane deep in the linear region cycles from the empty guess exactly as the restated Newton loop does (nmosfet.rs:51-102 returns
    # zero conductances for v_ds < 0): NOT_CONVERGED on that lane only
    sweep = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("VG"))
    out, status = sweep.solve_dc_sweep(np.array([1.0, 3.0, 1.5]), cfg=kb.SolverConfig(maximum_iterations=40))
    assert status.tolist() == [0, 2, 0]
