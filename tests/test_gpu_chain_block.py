"""GPU tier: the block-partitioned chain kernel (chain_block_kernel.cuh: one CTA per lane, Wang partition + parallel cyclic
reduction, results only through HBM) against the CPU oracle and against the sequential chain kernel on the same inputs.

Tolerance: 1e-9 max relative error vs the oracle; the two kernels stamp bit-identical values and differ only
in elimination order, so they are held to 5e-10 against each other (the host admits a plan to the block kernel on that bound).
"""
import numpy as np
import pytest

import netgen
from helpers import assert_close
from vm_emulator import chain_block_info

import krets_b200 as kb
from krets_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert kb.device_count() > 0, "GPU tier needs a CUDA device (the library has no CPU fallback)"


def _dc_point(oc, v):
    return oc.solve_dc("V1", v, v, 1.0)[0]


@pytest.mark.parametrize("nodes,shape", [(5000, (256, 21)), (3000, (256, 13)), (700, (256, 5))])
def test_block_vs_sequential_and_oracle_dc(oracle, monkeypatch, nodes, shape):
    import torch
    text = netgen.resistor_ladder(nodes)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0)
    valid, T, M, _, _, why = chain_block_info(plan)
    assert valid and (T, M) == shape, why
    lanes = 1237                                     # more lanes than resident CTAs: the persistent loop wraps
    vals = torch.linspace(-4.0, 5.0, lanes, dtype=torch.float64, device="cuda")
    before = kb.kernel_launch_count()
    out, status = plan.solve_dc_sweep(vals)
    torch.cuda.synchronize()
    assert kb.kernel_launch_count() > before and int(status.abs().max().item()) == 0
    monkeypatch.setenv("KRB_CHAIN_SEQ", "1")
    seq, status = plan.solve_dc_sweep(vals)
    torch.cuda.synchronize()
    monkeypatch.delenv("KRB_CHAIN_SEQ")
    assert int(status.abs().max().item()) == 0
    a, b = out.cpu().numpy(), seq.cpu().numpy()
    assert_close(a, b, rtol=5e-10, what="block vs sequential chain kernel")
    assert not np.array_equal(a, b) or nodes < 1000   # different elimination order: the block kernel really ran
    v = vals.cpu().numpy()
    for l in (0, 1, 618, lanes - 1):
        assert_close(a[l], _dc_point(oc, v[l]), what=f"block chain DC lane {l}")


@pytest.mark.parametrize("nodes,shape", [(5000, (512, 10)), (2500, (512, 6)), (2500, (256, 10))])
def test_block_vs_sequential_and_oracle_ac(oracle, monkeypatch, nodes, shape):
    import torch
    text = netgen.rc_ladder(nodes)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    if shape[0] == 256:                              # profiling shape: complex plans on 256-thread CTAs, two per SM
        monkeypatch.setenv("KRB_CHAIN_BLOCK_CT256", "1")
    plan = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e4)
    monkeypatch.delenv("KRB_CHAIN_BLOCK_CT256", raising=False)
    valid, T, M, _, _, why = chain_block_info(plan)
    assert valid and (T, M) == shape, why
    freqs = np.logspace(0, 8, 411)
    f_dev = torch.from_numpy(freqs).cuda()
    out, status = plan.solve_ac(f_dev)
    torch.cuda.synchronize()
    assert int(status.abs().max().item()) == 0
    monkeypatch.setenv("KRB_CHAIN_SEQ", "1")
    seq, _ = plan.solve_ac(f_dev)
    torch.cuda.synchronize()
    monkeypatch.delenv("KRB_CHAIN_SEQ")
    a, b = out.cpu().numpy(), seq.cpu().numpy()
    assert_close(a, b, rtol=5e-10, what="block vs sequential chain kernel (AC)")
    idx = [0, 100, 205, 410]
    assert_close(a[idx], oc.solve_ac(freqs[idx])[1], what="block chain AC sampled lanes")


def test_block_host_buffers_and_chunks(oracle):
    """KRB_MEM_HOST path: numpy in, numpy out, several chunks through the two-stream delivery pipeline."""
    text = netgen.resistor_ladder(5000)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0)
    assert chain_block_info(plan)[0]
    vals = np.linspace(0.0, 2.0, 301)
    out, status = plan.solve_dc_sweep(vals)
    assert not status.any()
    x1 = oc.solve_op()
    assert_close(out[1:], vals[1:, None] * x1[None, :], what="block chain, host buffers: linearity against one oracle solve")
    assert np.abs(out[0]).max() == 0.0


def test_block_small_forced_monte_carlo_order_and_status(oracle, monkeypatch):
    """KRB_CHAIN_BLOCK=1 admits small systems: Monte-Carlo term runs on every row, a pivot order that differs from the unknown
    order, and a per-lane zero pivot."""
    monkeypatch.setenv("KRB_CHAIN_BLOCK", "1")
    text = netgen.rc_ladder(30)
    circuit = kb.parse_circuit_description(text)
    varied = [circuit.element_index(f"R{k}") for k in range(1, 31)] + [circuit.element_index(f"C{k}") for k in range(1, 31)]
    lanes = 700
    rng = np.random.default_rng(11)
    vals = np.array([[circuit.element_value(e) * (0.95 + 0.1 * rng.random()) for _ in range(lanes)] for e in varied])
    freqs = np.logspace(1, 7, lanes)
    plan = kb.Plan(circuit, api.MODE_AC, varied=varied)
    assert plan.info.executor == 1 and chain_block_info(plan)[0]
    out, status = plan.solve_ac(freqs, varied_values=np.ascontiguousarray(vals))
    assert not status.any()
    for l in (0, 333, 699):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        assert_close(out[l], oc.solve_ac([freqs[l]])[1][0], what=f"block chain MC AC lane {l}")
    # resistors listed from the grounded end, source last: unknown order is the reverse of the elimination order
    lines = [f"R{k} n{k} n{k + 1} 1k" for k in range(30, 0, -1)] + ["R31 n31 0 1k", "V1 n1 0 2"]
    text = "\n".join(lines) + "\n"
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=circuit.element_index("V1"))
    assert plan.info.executor == 1 and chain_block_info(plan)[0]
    sv = np.array([0.5, 2.0, -7.0])
    out, status = plan.solve_dc_sweep(sv)
    assert not status.any()
    for l, v in enumerate(sv):
        assert_close(out[l], _dc_point(oc, v), what="block chain, permuted order")
    # a lane whose series resistor is +inf: g = 0 isolates the tail of the ladder -> exactly zero pivot on that lane only
    text = netgen.resistor_ladder(20)
    circuit = kb.parse_circuit_description(text)
    e = [circuit.element_index("R7"), circuit.element_index("R8"), circuit.element_index("R20")]
    plan = kb.Plan(circuit, api.MODE_DC, varied=e)
    assert plan.info.executor == 1 and chain_block_info(plan)[0]
    vv = np.array([[1000.0, np.inf], [1000.0, np.inf], [1000.0, np.inf]])
    out, status = plan.solve_op(2, varied_values=vv)
    assert status[0] == 0 and (status[1] & api.LANE_ZERO_PIVOT)
    assert_close(out[0], oracle.Circuit(text).solve_op(), what="healthy lane next to a singular one")


def test_block_shared_dictionary_many_generic_entries(oracle, monkeypatch):
    """~200 GENERIC entries in shared-dictionary mode on the 256-thread shape: more entries than the T/2 owner threads of the
    upper warps, so an owner rebuilds several entries per lane (chain_block_kernel.cuh: refresh)."""
    monkeypatch.setenv("KRB_CHAIN_BLOCK", "1")
    nodes = 200
    text = netgen.resistor_ladder(nodes)
    circuit = kb.parse_circuit_description(text)
    varied = [circuit.element_index(f"R{k}") for k in range(1, nodes + 1)]
    lanes = 333                                      # more lanes than resident CTAs
    rng = np.random.default_rng(5)
    vals = np.array([[circuit.element_value(e) * (0.9 + 0.2 * rng.random()) for _ in range(lanes)] for e in varied])
    plan = kb.Plan(circuit, api.MODE_DC, varied=varied)
    valid, T, M, n_dict, _, why = chain_block_info(plan)
    assert plan.info.executor == 1 and valid and T == 256, why
    out, status = plan.solve_op(lanes, varied_values=np.ascontiguousarray(vals))
    assert not status.any()
    monkeypatch.setenv("KRB_CHAIN_SEQ", "1")
    seq, _ = plan.solve_op(lanes, varied_values=np.ascontiguousarray(vals))
    monkeypatch.delenv("KRB_CHAIN_SEQ")
    assert_close(out, seq, rtol=5e-10, what="block vs sequential chain kernel, 200 generic entries")
    for l in (0, 150, lanes - 1):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        assert_close(out[l], oc.solve_op(), what=f"block chain, 200 generic entries, lane {l}")


@pytest.mark.parametrize("cplx", [False, True])
def test_block_prestamped_rows_monte_carlo(oracle, cplx):
    """Plans whose per-lane rows do not fit the shared dictionary (every element Monte-Carlo): the rows are stamped by
    krb_block_stamp_rows (thread per (row, lane)) and streamed by the block kernel.  More lanes than resident CTAs, a lane count
    that is not a multiple of the stamping tile, every lane against the oracle on a sample + block vs sequential kernel on all."""
    import torch
    nodes = 700
    text = netgen.rc_ladder(nodes) if cplx else netgen.resistor_ladder(nodes)
    circuit = kb.parse_circuit_description(text)
    varied = [e for e in range(circuit.num_elements) if circuit.element(e).kind in (api.KIND_R, api.KIND_C)]
    lanes = 1301
    rng = np.random.default_rng(23)
    vals = np.array([[circuit.element_value(e) * (0.95 + 0.1 * rng.random()) for _ in range(lanes)] for e in varied])
    plan = kb.Plan(circuit, api.MODE_AC if cplx else api.MODE_DC, varied=varied, nominal_frequency=1e4)
    valid, T, M, n_dict, _, why = chain_block_info(plan)
    assert plan.info.executor == 1 and valid, why
    v_dev = torch.from_numpy(vals).cuda()
    if cplx:
        freqs = np.logspace(1, 7, lanes)
        out, status = plan.solve_ac(torch.from_numpy(freqs).cuda(), varied_values=v_dev)
    else:
        out, status = plan.solve_op(lanes, varied_values=v_dev)
    torch.cuda.synchronize()
    assert int(status.abs().max().item()) == 0
    a = out.cpu().numpy()
    import os
    os.environ["KRB_CHAIN_SEQ"] = "1"
    try:
        seq = (plan.solve_ac(torch.from_numpy(freqs).cuda(), varied_values=v_dev) if cplx else plan.solve_op(lanes, varied_values=v_dev))[0].cpu().numpy()
    finally:
        del os.environ["KRB_CHAIN_SEQ"]
    assert_close(a, seq, rtol=5e-10, what="pre-stamped block kernel vs sequential chain kernel")
    assert not np.array_equal(a, seq)
    for l in (0, 1, 650, lanes - 1):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        ref = oc.solve_ac([freqs[l]])[1][0] if cplx else oc.solve_op()
        assert_close(a[l], ref, what=f"pre-stamped block kernel lane {l}")


@pytest.mark.parametrize("cplx", [False, True])
def test_block_prestamped_rows_read_device_values_in_place(oracle, cplx):
    """Device-resident Monte-Carlo values that already have the kernels' layout (one chunk, lane count a multiple of the block size) are
    read in place by krb_block_stamp_rows instead of being copied into the plan's input planes: same bits as the copying path (host
    arrays), and the caller's array is left untouched."""
    import torch
    nodes = 600
    text = netgen.rc_ladder(nodes) if cplx else netgen.resistor_ladder(nodes)
    circuit = kb.parse_circuit_description(text)
    varied = [e for e in range(circuit.num_elements) if circuit.element(e).kind in (api.KIND_R, api.KIND_C)]
    lanes = 1344                                   # 21 x 64: no lane padding
    rng = np.random.default_rng(5)
    vals = np.array([[circuit.element_value(e) * (0.95 + 0.1 * rng.random()) for _ in range(lanes)] for e in varied])
    plan = kb.Plan(circuit, api.MODE_AC if cplx else api.MODE_DC, varied=varied, nominal_frequency=1e4)
    assert plan.info.executor == 1 and chain_block_info(plan)[0]
    v_dev = torch.from_numpy(vals).cuda()
    freqs = np.logspace(1, 7, lanes)
    if cplx:
        dev = plan.solve_ac(torch.from_numpy(freqs).cuda(), varied_values=v_dev)
        host = plan.solve_ac(freqs, varied_values=vals)
    else:
        dev = plan.solve_op(lanes, varied_values=v_dev)
        host = plan.solve_op(lanes, varied_values=vals)
    torch.cuda.synchronize()
    assert int(dev[1].abs().max().item()) == 0 and not host[1].any()
    assert np.array_equal(dev[0].cpu().numpy(), host[0])
    assert np.array_equal(v_dev.cpu().numpy(), vals)
    for l in (0, 777, lanes - 1):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        ref = oc.solve_ac([freqs[l]])[1][0] if cplx else oc.solve_op()
        assert_close(host[0][l], ref, what=f"lane {l}")
