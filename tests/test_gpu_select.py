"""GPU tier: result selection (krb_io.select — a labelled EXTENSION: the reference returns every unknown of every point,
ac.rs:97-105).  The selected columns must be bit-identical to the same columns of the full delivery, on every executor and on
both memory sides."""
import numpy as np
import pytest

import netgen

import krets_b200 as kb
from krets_b200 import api

pytestmark = pytest.mark.gpu


def _same(full, part, sel):
    assert part.shape[-1] == len(sel)
    assert np.array_equal(np.asarray(full)[..., sel], np.asarray(part))


@pytest.mark.parametrize("nodes,env", [(5000, {}), (300, {}), (300, {"KRB_CHAIN_BLOCK": "1"})])
def test_select_chain_executors(monkeypatch, nodes, env):
    import torch
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    circuit = kb.parse_circuit_description(netgen.resistor_ladder(nodes))
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0)
    for k in env:
        monkeypatch.delenv(k)
    sel = [circuit.index_map["I(V1)"], circuit.index_map[f"V({nodes})"], circuit.index_map["V(2)"], circuit.index_map[f"V({nodes // 2})"]]
    vals = np.linspace(-2.0, 3.0, 777)
    full, st0 = plan.solve_dc_sweep(vals)
    part, st1 = plan.solve_dc_sweep(vals, select=sel)
    assert not st0.any() and not st1.any()
    _same(full, part, sel)
    part_d, _ = plan.solve_dc_sweep(torch.from_numpy(vals).cuda(), select=sel)
    _same(full, part_d.cpu().numpy(), sel)


def test_select_ac_vm_transient_and_front():
    circuit = kb.parse_circuit_description(netgen.rc_ladder(600))
    plan = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e4)
    freqs = np.logspace(0, 8, 333)
    sel = [5, 600, 1]
    full, _ = plan.solve_ac(freqs)
    part, _ = plan.solve_ac(freqs, select=sel)
    _same(full, part, sel)
    # lane VM (forced) and the dense transient kernel
    text = "V1 in 0 dc 0 PULSE(0 1 1u 1u 1u 1 1)\nR1 in mid 1k\nC1 mid 0 1u\nR2 mid out 2k\nC2 out 0 100n\n"
    circuit = kb.parse_circuit_description(text)
    pdc, ptr = kb.Plan(circuit, api.MODE_DC), kb.Plan(circuit, api.MODE_TRAN, time_step=1e-5)
    sel = [circuit.index_map["V(out)"], circuit.index_map["I(V1)"]]
    full, _ = ptr.solve_transient(pdc, 200, 3)
    part, _ = ptr.solve_transient(pdc, 200, 3, select=sel)
    _same(full, part, sel)
    plan = kb.Plan(circuit, api.MODE_DC, force_lane_vm=True, sweep_element=0)
    vals = np.linspace(0, 1, 50)
    _same(plan.solve_dc_sweep(vals)[0], plan.solve_dc_sweep(vals, select=sel)[0], sel)
    # front executor
    circuit = kb.parse_circuit_description(netgen.rlc_mesh(9, 8, transient=True))
    plan = kb.Plan(circuit, api.MODE_DC, executor=api.EXEC_STAGED, sweep_element=circuit.element_index("I1"))
    vals = np.linspace(0, 2e-3, 11)
    sel = [0, circuit.n - 1, 7]
    _same(plan.solve_dc_sweep(vals)[0], plan.solve_dc_sweep(vals, select=sel)[0], sel)


def test_select_rejects_bad_arguments():
    circuit = kb.parse_circuit_description(netgen.resistor_ladder(30))
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0)
    with pytest.raises(kb.KretsError):
        plan.solve_dc_sweep(np.array([1.0]), select=[99])
    with pytest.raises(kb.KretsError):
        plan.solve_dc_sweep(np.array([1.0]), select=[1], layout=api.LAYOUT_LANE_INTERLEAVED, out=np.zeros((1, 1)))
