"""CPU tier: host symbolic analysis of the front executor (krets_b200/csrc/fronts.cpp).  The plan tables are run through a
numpy multifrontal emulator (tests/front_emulator.py) on matrices stamped by the oracle and compared with SuperLU, which
checks supernodes, relative indices, assembly lists and level order without a GPU.  No compute goes through the product."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

import front_emulator as fe
import netgen

import krets_b200 as kb
from krets_b200 import api


def _check_tables(ft, n):
    piv0, idx_ptr, idx, parent = ft["piv0"], ft["idx_ptr"], ft["idx"], ft["parent"]
    nf = ft["n_fronts"]
    assert piv0[0] == 0 and piv0[-1] == n and np.all(np.diff(piv0) > 0)
    level = np.zeros(nf, dtype=int)
    for l in range(ft["n_levels"]):
        level[ft["level_fronts"][ft["level_ptr"][l]:ft["level_ptr"][l + 1]]] = l
    assert sorted(ft["level_fronts"].tolist()) == list(range(nf))
    for f in range(nf):
        upd = idx[idx_ptr[f]:idx_ptr[f + 1]]
        assert np.all(np.diff(upd) > 0) and (len(upd) == 0 or upd[0] >= piv0[f + 1])
        if parent[f] >= 0:
            assert parent[f] > f and level[parent[f]] > level[f]
            assert piv0[parent[f]] <= upd[0] < piv0[parent[f] + 1]       # first update index is a pivot of the parent
        else:
            assert len(upd) == 0


@pytest.mark.parametrize("shape,transient", [((5, 6), True), ((12, 13), True), ((12, 13), False), ((24, 20), True)])
@pytest.mark.parametrize("mode", ["dc", "ac", "tran"])
def test_front_plan_emulated(oracle, shape, transient, mode):
    if mode == "ac" and transient:
        pytest.skip("AC with a current source is structurally singular in the reference (SURVEY §B-3)")
    text = netgen.rlc_mesh(*shape, transient=transient)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    kmode = {"dc": api.MODE_DC, "ac": api.MODE_AC, "tran": api.MODE_TRAN}[mode]
    par = {"dc": 0.0, "ac": 2.5e6, "tran": 1e-9}[mode]
    plan = kb.Plan(circuit, kmode, executor=api.EXEC_STAGED, time_step=1e-9 if mode == "tran" else 0.0, nominal_frequency=1e6)
    assert plan.info.executor == 2
    ft = plan.front_tables()
    _check_tables(ft, circuit.n)
    colptr, rowidx = plan.pattern()
    ocp, ori, vals = oc.stamp_matrix(mode, par)
    assert np.array_equal(colptr, ocp) and np.array_equal(rowidx[:len(ori)], ori)      # bit-exact stamped pattern
    ro, co = plan.pivot_order()
    assert sorted(ro.tolist()) == list(range(circuit.n)) and sorted(co.tolist()) == list(range(circuit.n))
    rng = np.random.default_rng(7)
    cplx = mode == "ac"
    b = rng.standard_normal(circuit.n) + (1j * rng.standard_normal(circuit.n) if cplx else 0.0)
    x = fe.factor_and_solve(ft, ro, co, vals, b, dtype=np.complex128 if cplx else np.float64)
    A = sp.csc_matrix((vals, ori, ocp), shape=(circuit.n, circuit.n))
    xr = spl.splu(A).solve(b)
    assert np.max(np.abs(x - xr)) <= 1e-10 * np.max(np.abs(xr))
    assert plan.info.nnz_lu == ft["lu_size"] or plan.info.nnz_lu <= ft["lu_size"]


def test_front_plan_nominal_dc_needs_in_front_pivoting(oracle):
    """DC mesh with identical element values: inductors are shorts (no branch diagonal, stampable.rs:688-701) and a static
    pivot cancels exactly; the per-lane threshold pivoting inside the fully-summed block recovers full accuracy."""
    text = netgen.rlc_mesh(40, 40, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(circuit, api.MODE_DC, executor=api.EXEC_STAGED)
    ocp, ori, vals = oc.stamp_matrix("dc", 0.0)
    ro, co = plan.pivot_order()
    b = np.zeros(circuit.n)
    b[circuit.index_map["I(I1)"]] = 1e-3
    with np.errstate(all="ignore"):
        x = fe.factor_and_solve(plan.front_tables(), ro, co, vals, b)
    xr = spl.splu(sp.csc_matrix((vals, ori, ocp), shape=(circuit.n, circuit.n))).solve(b)
    assert np.max(np.abs(x - xr)) <= 1e-10 * np.max(np.abs(xr))


def test_front_executor_selection():
    small = kb.parse_circuit_description(netgen.rlc_mesh(4, 4))
    assert kb.Plan(small, api.MODE_AC).info.executor == 0                      # small fill: per-lane program
    big = kb.parse_circuit_description(netgen.rlc_mesh(30, 30))
    assert kb.Plan(big, api.MODE_AC).info.executor == 2                        # mesh: front executor
    assert kb.Plan(big, api.MODE_AC, force_lane_vm=True).info.executor == 0
    ladder = kb.parse_circuit_description(netgen.resistor_ladder(300))
    assert kb.Plan(ladder, api.MODE_DC, sweep_element=0).info.executor == 1     # ladders keep the chain kernel
    assert kb.Plan(ladder, api.MODE_DC, sweep_element=0, executor=api.EXEC_STAGED).info.executor == 2


def test_front_rejects_nonlinear(golden):
    circuit = kb.parse_circuit_description(golden["netlists"]["diode_iv_curve"])
    with pytest.raises(kb.KretsError) as ei:
        kb.Plan(circuit, api.MODE_DC, executor=api.EXEC_STAGED)
    assert ei.value.code == 7


def test_front_needs_a_device():
    if kb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    circuit = kb.parse_circuit_description(netgen.rlc_mesh(5, 6, transient=True))
    plan = kb.Plan(circuit, api.MODE_DC, executor=api.EXEC_STAGED)
    with pytest.raises(kb.KretsError) as ei:
        plan.solve_op(1)
    assert ei.value.code == 20      # KRB_ERR_NO_DEVICE: there is no CPU fallback
