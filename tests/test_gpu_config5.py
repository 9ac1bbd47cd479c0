"""GPU tier: BASELINE config 5 at its stated size — the 224 x 224 RLC mesh (n ~ 100 130), front executor, Monte-Carlo lanes.

The oracle's natural-order LU cannot factor this system in seconds, so parity is held through
  * the residual of every checked point on the ORACLE-stamped matrix (oracle.Circuit.stamp_matrix), ||A x - b||_inf / ||b||_inf <= 1e-9,
  * agreement with SuperLU (scipy, own ordering + partial pivoting) on that same matrix to 1e-8,
  * for the transient, agreement of all 9 rows with tests/mesh_reference.trap_rows (itself held to the oracle's trapezoidal driver
    on a small mesh by tests/test_mesh_reference_cpu.py) to 1e-8.
AC lanes cover 10 kHz, 1 MHz and 100 MHz with the static pivots planned at 10 kHz, as bench.py's mesh_ac workload does.
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spl

import mesh_reference as mr
import netgen
from helpers import assert_close

import krets_b200 as kb
from krets_b200 import api

pytestmark = pytest.mark.gpu
GRID = 224


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert kb.device_count() > 0, "GPU tier needs a CUDA device (the library has no CPU fallback)"


def _variants(circuit, oc, tab, lanes):
    """per-lane element values, SURVEY 8(d) row 5 rule (splitmix64, element ordinal as index, lane 0 nominal)"""
    varied = np.nonzero((tab.kind == mr.KIND_R) | (tab.kind == mr.KIND_C) | (tab.kind == mr.KIND_L))[0]
    nominal = tab.values(oc)
    vals = np.empty((len(varied), lanes))
    for lane in range(lanes):
        vals[:, lane] = [nominal[e] * netgen.mc_factor(lane, int(e), circuit.num_elements) for e in varied]
    return [int(e) for e in varied], nominal, vals


def _set_lane(oc, varied, vals, lane):
    for k, e in enumerate(varied):
        oc.set_element_value(e, vals[k, lane])


def test_config5_ac_full_size(oracle):
    text = netgen.rlc_mesh(GRID, GRID, transient=False)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    assert circuit.n > 100000
    tab = mr.ElementTable(oc)
    freqs = np.array([1e4, 1e6, 1e8, 1e4, 1e6, 1e8])
    lanes = len(freqs)
    varied, nominal, vals = _variants(circuit, oc, tab, lanes)
    plan = kb.Plan(circuit, api.MODE_AC, varied=varied, nominal_frequency=1e4, executor=api.EXEC_STAGED)
    assert plan.info.executor == 2
    out, status = plan.solve_ac(freqs, varied_values=vals)
    assert not status.any(), status
    for lane in range(lanes):
        _set_lane(oc, varied, vals, lane)
        A = mr.csc(oc, "ac", float(freqs[lane]))
        b = mr.ac_rhs(oc, tab, tab.values(oc))
        res = np.max(np.abs(A @ out[lane] - b)) / np.max(np.abs(b))
        assert res <= 1e-9, f"lane {lane} ({freqs[lane]:g} Hz): residual {res:.2e} on the oracle-stamped matrix"
        assert_close(out[lane], spl.splu(A).solve(b), rtol=1e-8, what=f"224x224 AC lane {lane} ({freqs[lane]:g} Hz) vs SuperLU")


def test_config5_transient_full_size(oracle):
    text = netgen.rlc_mesh(GRID, GRID, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    tab = mr.ElementTable(oc)
    lanes, h, steps = 3, 1e-9, 8
    varied, nominal, vals = _variants(circuit, oc, tab, lanes)
    pdc = kb.Plan(circuit, api.MODE_DC, varied=varied, executor=api.EXEC_STAGED)
    ptr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h, integrator=api.INTEGRATOR_TRAP, executor=api.EXEC_STAGED)
    out, status = ptr.solve_transient(pdc, steps, lanes, varied_values=vals)
    assert not status.any() and out.shape == (lanes, steps + 1, circuit.n)
    for lane in range(lanes):
        _set_lane(oc, varied, vals, lane)
        # row 0: the operating point on the oracle-stamped DC matrix
        A0 = mr.csc(oc, "dc", 0.0)
        b0 = mr.dc_rhs(oc, tab, tab.values(oc))
        res = np.max(np.abs(A0 @ out[lane, 0] - b0)) / np.max(np.abs(b0))
        assert res <= 1e-9, f"lane {lane}: operating-point residual {res:.2e}"
        ref = mr.trap_rows(oc, tab, h, steps)
        assert_close(out[lane], ref, rtol=1e-8, what=f"224x224 transient lane {lane}: OP + {steps} trapezoidal steps vs SuperLU stepping")
