"""GPU tier (-m gpu): the supernodal front executor (KRB_EXEC_STAGED, krets_b200/csrc/front_kernels.cuh) against the CPU
oracle on RLC meshes (SURVEY §8d config 5 generator at reduced size), through the C ABI.  Tolerance: max relative error
1e-9 per unknown (helpers.assert_close).  At sizes the oracle's natural-order LU cannot handle in seconds, the check is the
size-independent one: the residual of the oracle-stamped matrix and agreement with SuperLU (scipy) on the same matrix."""
import os

import numpy as np
import pytest

import netgen
from helpers import assert_close

import krets_b200 as kb
from krets_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert kb.device_count() > 0, "GPU tier needs a CUDA device (the library has no CPU fallback)"


def mc_values(circuit, oc, lanes):
    """±5 % Monte-Carlo on every R/L/C (splitmix64 rule, element ordinal as index; lane 0 nominal)."""
    varied = [e for e in range(circuit.num_elements) if oc.element_kind(e) in (api.KIND_R, api.KIND_C, api.KIND_L)]
    vals = np.zeros((len(varied), lanes))
    for k, e in enumerate(varied):
        nominal = oc.element_value(e)
        for lane in range(lanes):
            vals[k, lane] = nominal * netgen.mc_factor(lane, e, circuit.num_elements)
    return varied, vals


def oracle_lane(oc, varied, vals, lane, fn):
    saved = [oc.element_value(e) for e in varied]
    for k, e in enumerate(varied):
        oc.set_element_value(e, vals[k, lane])
    try:
        return fn()
    finally:
        for k, e in enumerate(varied):
            oc.set_element_value(e, saved[k])


@pytest.mark.parametrize("shape", [(5, 6), (12, 13)])
def test_front_op_monte_carlo(oracle, shape):
    text = netgen.rlc_mesh(*shape, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    lanes = 6
    varied, vals = mc_values(circuit, oc, lanes)
    plan = kb.Plan(circuit, api.MODE_DC, varied=varied, executor=api.EXEC_STAGED)
    assert plan.info.executor == 2
    before = kb.kernel_launch_count()
    out, status = plan.solve_op(lanes, varied_values=vals)
    assert kb.kernel_launch_count() > before and not status.any()
    for lane in range(lanes):
        ref = oracle_lane(oc, varied, vals, lane, oc.solve_op)
        assert_close(out[lane], ref, what=f"mesh {shape} OP lane {lane}")


def test_front_dc_sweep(oracle):
    text = netgen.rlc_mesh(9, 8, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=circuit.element_index("I1"), executor=api.EXEC_STAGED)
    vals = np.linspace(0.0, 2e-3, 37)
    out, status = plan.solve_dc_sweep(vals)
    ref = oc.solve_dc("I1", 0.0, 2e-3, 2e-3 / 36)
    assert not status.any() and out.shape[0] == 37
    assert_close(out[: ref.shape[0]], ref[:37], what="mesh DC sweep")


@pytest.mark.parametrize("shape", [(5, 6), (12, 13)])
def test_front_ac(oracle, shape):
    text = netgen.rlc_mesh(*shape, transient=False)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = kb.generate_frequencies("Decade", 5, 1.0, 1e9)
    plan = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e6, executor=api.EXEC_STAGED)
    out, status = plan.solve_ac(freqs)
    _, ref = oc.solve_ac(freqs)
    assert not status.any()
    assert_close(out, ref, what=f"mesh {shape} AC")


@pytest.mark.parametrize("integrator", [api.INTEGRATOR_BE_KRETS, api.INTEGRATOR_TRAP, api.INTEGRATOR_BE])
def test_front_transient_monte_carlo(oracle, integrator):
    text = netgen.rlc_mesh(6, 7, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    lanes, h, steps = 3, 2e-7, 150
    varied, vals = mc_values(circuit, oc, lanes)
    plan_dc = kb.Plan(circuit, api.MODE_DC, varied=varied, executor=api.EXEC_STAGED)
    plan_tr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h, integrator=integrator, executor=api.EXEC_STAGED)
    out, status = plan_tr.solve_transient(plan_dc, steps, lanes, varied_values=vals)
    assert not status.any() and out.shape == (lanes, steps + 1, circuit.n)
    for lane in range(lanes):
        _, ref = oracle_lane(oc, varied, vals, lane, lambda: oc.solve_tran(h, steps * h, integrator=integrator))
        assert_close(out[lane], ref, rtol=1e-8, what=f"mesh transient integrator {integrator} lane {lane}")


@pytest.mark.parametrize("ll", ["1", "0"])
def test_front_blocked_path_matches(oracle, monkeypatch, ll):
    """Fronts beyond shared memory: left-looking by resident column blocks (krb_front_factor_ll, ll = 1) or, with KRB_FRONT_LL=0,
    the blocked right-looking kernel on the per-lane workspace; forcing every front larger than 6 onto them must give the same answers."""
    monkeypatch.setenv("KRB_FRONT_SMEM_M", "6")
    monkeypatch.setenv("KRB_FRONT_LL", ll)
    text = netgen.rlc_mesh(12, 13, transient=False)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = kb.generate_frequencies("Decade", 3, 1e3, 1e9)
    out, status = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e6, executor=api.EXEC_STAGED).solve_ac(freqs)
    _, ref = oc.solve_ac(freqs)
    assert not status.any()
    assert_close(out, ref, what="blocked path AC")
    text = netgen.rlc_mesh(12, 13, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    out, status = kb.Plan(circuit, api.MODE_DC, executor=api.EXEC_STAGED).solve_op(1)
    assert_close(out[0], oc.solve_op(), what="blocked path OP")


def test_front_mixed_kernel_routing(oracle, monkeypatch):
    """Fronts beyond the left-looking kernel's one-row-per-thread limit (512 rows; lowered to 40 here) fall back to the blocked kernel on
    the per-lane workspace or the CTA-per-front kernel, smaller ones stay on the left-looking / warp kernels: one factorisation mixes all
    of them and children's contribution blocks cross between the kernels."""
    monkeypatch.setenv("KRB_FRONT_LL_MAX_M", "40")
    monkeypatch.setenv("KRB_FRONT_SMEM_M", "56")
    text = netgen.rlc_mesh(30, 31, transient=False)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = kb.generate_frequencies("Decade", 2, 1e4, 1e8)
    out, status = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e6, executor=api.EXEC_STAGED).solve_ac(freqs)
    _, ref = oc.solve_ac(freqs)
    assert not status.any()
    assert_close(out, ref, what="mixed kernel routing AC")


@pytest.mark.parametrize("smem_m", [None, 24, -24])
def test_front_nominal_dc_pivots_inside_fronts(oracle, monkeypatch, smem_m):
    """Identical element values + inductor shorts: a static pivot cancels exactly (CPU tier shows it); the kernels' per-lane
    threshold pivoting inside the fully-summed block must recover it, on the shared-memory and on the blocked path."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    if smem_m:
        monkeypatch.setenv("KRB_FRONT_SMEM_M", str(abs(smem_m)))
        if smem_m < 0:                                      # the kernels the left-looking one replaced (CTA per front, workspace)
            monkeypatch.setenv("KRB_FRONT_LL", "0")
            monkeypatch.setenv("KRB_FRONT_WARPF", "0")
    text = netgen.rlc_mesh(40, 40, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    out, status = kb.Plan(circuit, api.MODE_DC, executor=api.EXEC_STAGED).solve_op(1)
    assert not status.any()
    colptr, rowidx, v = oc.stamp_matrix("dc", 0.0)
    A = sp.csc_matrix((v, rowidx, colptr), shape=(circuit.n, circuit.n))
    b = np.zeros(circuit.n)
    b[circuit.index_map["I(I1)"]] = 1e-3
    assert_close(out[0], spl.splu(A).solve(b), rtol=1e-8, what="40x40 nominal DC vs SuperLU")


def test_front_chunks_and_windows(oracle, monkeypatch):
    """Lane chunks (resident factor stores) and factor windows (workspace) that do not divide the lane count."""
    monkeypatch.setenv("KRB_FRONT_CHUNK", "3")
    monkeypatch.setenv("KRB_FRONT_WINDOW", "2")
    text = netgen.rlc_mesh(6, 7, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    lanes, h, steps = 7, 2e-7, 40
    varied, vals = mc_values(circuit, oc, lanes)
    plan_dc = kb.Plan(circuit, api.MODE_DC, varied=varied, executor=api.EXEC_STAGED)
    plan_tr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h, integrator=api.INTEGRATOR_TRAP, executor=api.EXEC_STAGED)
    out, status = plan_tr.solve_transient(plan_dc, steps, lanes, varied_values=vals)
    assert not status.any()
    import torch
    out_d, status_d = plan_tr.solve_transient(plan_dc, steps, lanes, varied_values=torch.from_numpy(vals).cuda())
    assert np.array_equal(out_d.cpu().numpy(), out) and not status_d.any().item()
    for lane in (0, 2, 3, 6):
        _, ref = oracle_lane(oc, varied, vals, lane, lambda: oc.solve_tran(h, steps * h, integrator=api.INTEGRATOR_TRAP))
        assert_close(out[lane], ref, rtol=1e-8, what=f"chunked transient lane {lane}")
    text = netgen.rlc_mesh(6, 7, transient=False)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = kb.generate_frequencies("Decade", 1, 1e3, 1e9)
    out, status = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e3, executor=api.EXEC_STAGED).solve_ac(freqs)
    _, ref = oc.solve_ac(freqs)
    assert not status.any() and len(freqs) == 7
    assert_close(out, ref, what="chunked AC")


def test_front_zero_pivot_is_reported():
    text = netgen.rlc_mesh(5, 6, transient=True)
    circuit = kb.parse_circuit_description(text)
    e = circuit.element_index("Rs")
    plan = kb.Plan(circuit, api.MODE_DC, varied=[e], executor=api.EXEC_STAGED)
    out, status = plan.solve_op(2, varied_values=np.array([[1.0, 0.0]]))   # lane 1: 1/0 conductance -> non-finite pivot
    assert status[0] == 0 and (status[1] & api.LANE_ZERO_PIVOT)


def test_front_selected_automatically_and_solver_api(oracle):
    text = netgen.rlc_mesh(30, 30, transient=False)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    assert kb.Plan(circuit, api.MODE_AC).info.executor == 2          # too much fill for a per-lane program
    res = kb.Solver(circuit).solve(kb.Analysis.Ac(kb.AcAnalysis(kb.AcSweep.Decade(1), 1e3, 1e6)))
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    for row, f in zip(res.array, res.axis):
        colptr, rowidx, vals = oc.stamp_matrix("ac", float(f))
        A = sp.csc_matrix((vals, rowidx, colptr), shape=(circuit.n, circuit.n))
        b = np.zeros(circuit.n, dtype=complex)
        b[circuit.index_map["I(V1)"]] = 1.0
        assert np.max(np.abs(A @ row - b)) < 1e-9
        assert_close(row, spl.splu(A).solve(b), rtol=1e-8, what=f"30x30 mesh AC at {f} Hz vs SuperLU")


def test_front_medium_mesh_residual(oracle):
    """64 x 64 mesh (n = 8131): transient companion matrix at h = 1 ns, Monte-Carlo lanes; SuperLU on the oracle-stamped
    matrix is the reference, plus the residual (size-independent property)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    text = netgen.rlc_mesh(64, 64, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    lanes = 4
    varied, vals = mc_values(circuit, oc, lanes)
    plan = kb.Plan(circuit, api.MODE_DC, varied=varied, executor=api.EXEC_STAGED)
    out, status = plan.solve_op(lanes, varied_values=vals)
    assert not status.any()
    for lane in range(lanes):
        def ref():
            colptr, rowidx, v = oc.stamp_matrix("dc", 0.0)
            A = sp.csc_matrix((v, rowidx, colptr), shape=(circuit.n, circuit.n))
            b = np.zeros(circuit.n)
            b[circuit.index_map["I(I1)"]] = 1e-3
            return A, b, spl.splu(A).solve(b)
        A, b, x = oracle_lane(oc, varied, vals, lane, ref)
        assert np.max(np.abs(A @ out[lane] - b)) < 1e-12
        assert_close(out[lane], x, rtol=1e-8, what=f"64x64 mesh OP lane {lane} vs SuperLU")
