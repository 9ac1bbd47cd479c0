"""CPU tier: the C-ABI library loads and exports what the headers declare; host-side logic (netlist front end,
index map, frequency grids, symbolic plan, lane-program compiler) checked against the oracle.  No GPU calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import netgen
from helpers import assert_close
from vm_emulator import Emulator, input_layout, run_sequential_sweep

import krets_b200 as kb
from krets_b200 import _lib, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kcirc(golden, key):
    if key.startswith("resistor_ladder_"):
        return kb.parse_circuit_description(netgen.resistor_ladder(int(key.rsplit("_", 1)[1])))
    return kb.parse_circuit_description(golden["netlists"][key])


def ocirc(oracle, golden, key):
    if key.startswith("resistor_ladder_"):
        return oracle.Circuit(netgen.resistor_ladder(int(key.rsplit("_", 1)[1])))
    return oracle.Circuit(golden["netlists"][key])


ALL = ["low_pass_filter", "low_pass_filter_transient", "dual_rc_ladder", "high_pass_filter", "high_pass_filter_transient",
       "voltage_divider", "basic_001", "diode_iv_curve", "rectifier", "farid_n_najm", "inverter", "test_case_1", "resistor_ladder_500"]


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = set()
    for hdr in ("krets_b200.h", "krets_b200_host.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        declared |= set(re.findall(r"\b(kr[bh]_[a-z0-9_]+)\s*\(", src))
    assert declared, "no declarations found"
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib.SYMBOLS) == declared            # the ctypes table binds exactly the public surface
    assert "sm_100a" in kb.version()


def test_no_device_fails_loudly(golden):
    """No CPU fallback: without a GPU every numeric entry point returns KRB_ERR_NO_DEVICE."""
    if kb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    c = kcirc(golden, "voltage_divider")
    with pytest.raises(kb.KretsError) as e:
        kb.Solver(c).solve(kb.Analysis.Op())
    assert e.value.kind == "NoDevice"


@pytest.mark.parametrize("key", ALL)
def test_index_map_bit_exact(oracle, golden, key):
    """parser.rs:105-125 — MNA unknown numbering must be identical to the reference's."""
    kc, oc = kcirc(golden, key), ocirc(oracle, golden, key)
    assert kc.names == oc.names
    assert kc.num_elements == oc.num_elements
    for e in range(kc.num_elements):
        el = kc.element(e)
        # NMOS: ic carries the source node (krets_b200.h); the index map has no I(M..) key, so the oracle's lookup is -1
        assert (el.ip, el.im, -1 if el.kind == api.KIND_NMOS else el.ic) == oc.element_indices(e)
        assert kc.identifiers[e] == oc.element_ident(e)
        assert el.kind == oc.element_kind(e)
        assert el.value == oc.element_value(e)


def test_parser_errors():
    for text, kind in [("* nothing\n", "EmptyNetlist"), ("R1 a 0 1.5x\n", "ParseError"), ("D1 a 0 NOPE\nR1 a 0 1\n", "MissingModel"),
                       ("Q1 a b c\n", "UnsupportedElement")]:
        with pytest.raises(kb.KretsError) as e:
            kb.parse_circuit_description(text)
        assert e.value.kind == kind


def test_frequency_grids_bit_exact(oracle):
    for kind, pts, f0, f1 in [("Linear", 10, 0.1, 1000.0), ("Linear", 100, 1.0, 1000.0), ("Decade", 10, 1.0, 1e6),
                              ("Decade", 1250, 1.0, 1e8), ("Octave", 7, 3.0, 5000.0), ("Linear", 1, 5.0, 9.0), ("Decade", 3, 0.0, 1.0)]:
        a, b = kb.generate_frequencies(kind, pts, f0, f1), oracle.generate_frequencies(kind, pts, f0, f1)
        assert a.tobytes() == b.tobytes(), (kind, pts)


MODES = {"dc": api.MODE_DC, "ac": api.MODE_AC, "tran": api.MODE_TRAN}


@pytest.mark.parametrize("key,mode", [("resistor_ladder_500", "dc"), ("farid_n_najm", "dc"), ("basic_001", "dc"), ("low_pass_filter", "ac"),
                                      ("high_pass_filter", "ac"), ("dual_rc_ladder", "tran"), ("high_pass_filter_transient", "tran"),
                                      ("rectifier", "tran"), ("diode_iv_curve", "dc"), ("inverter", "dc")])
def test_sparsity_pattern_bit_exact(oracle, golden, key, mode):
    """The stamped (row,col) set after sum_triplets must equal the reference's CSC pattern."""
    kc, oc = kcirc(golden, key), ocirc(oracle, golden, key)
    plan = kb.Plan(kc, MODES[mode], time_step=1e-5)
    colptr, rowidx = plan.pattern()
    ocp, ori, _ = oc.stamp_matrix(mode, 1e-5 if mode == "tran" else 1000.0)
    assert colptr.tolist() == ocp.tolist() and rowidx.tolist() == ori.tolist()
    r, c = plan.pivot_order()
    assert sorted(r.tolist()) == list(range(kc.n)) and sorted(c.tolist()) == list(range(kc.n))


def test_ladder_plan_is_fill_free():
    """the 5000 ladder has a zero-fill order: nnz(L+U) must be 15 000 (= nnz(A))."""
    kc = kb.parse_circuit_description(netgen.resistor_ladder(5000))
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=0)
    assert (plan.info.n, plan.info.nnz, plan.info.nnz_lu) == (5001, 15000, 15000)
    assert plan.info.alg_bytes_per_point == 8 + 8 * (4 * 15000 + 3 * 5001)      # 600 032 B, SURVEY §8(d)
    kc = kb.parse_circuit_description(netgen.rc_ladder(5000))
    plan = kb.Plan(kc, api.MODE_AC)
    assert (plan.info.n, plan.info.nnz, plan.info.nnz_lu) == (5001, 15000, 15000)
    assert plan.info.alg_bytes_per_point == 8 + 16 * (4 * 15000 + 3 * 5001)     # 1 200 056 B


def test_structural_singularity_and_unsupported(golden):
    # AC + current source -> empty row -> DecompositionFailed at analysis time
    with pytest.raises(kb.KretsError) as e:
        kb.Plan(kcirc(golden, "basic_001"), api.MODE_AC)
    assert e.value.kind == "DecompositionFailed"
    # NMOS AC is todo!() in the reference
    with pytest.raises(kb.KretsError) as e:
        kb.Plan(kcirc(golden, "inverter"), api.MODE_AC)
    assert e.value.kind == "Unsupported"
    with pytest.raises(kb.KretsError) as e:
        kb.Plan(kcirc(golden, "voltage_divider"), api.MODE_DC, sweep_element=1)   # R1 is not a source
    assert e.value.kind == "InvalidElementFormat"


# ---- the compiled lane program, emulated on the CPU, against the oracle ----------------------------------------
def test_program_dc_sweep_ladder(oracle):
    kc = kb.parse_circuit_description(netgen.resistor_ladder(500))
    oc = oracle.Circuit(netgen.resistor_ladder(500))
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=0, scratch_slots=16)
    vals = np.linspace(-2.0, 3.0, 7)
    em = Emulator(plan, len(vals))
    X = em.run(vals[None, :])
    ref = oc.solve_dc("V1", -2.0, 3.0, 5.0 / 6.0)
    assert ref.shape[0] == 7
    assert_close(X[0].T, ref, what="ladder DC sweep")
    assert not em.status.any()
    assert 0 < em.stats["stg"] <= em.stats["ldg"] <= 1.05 * em.stats["stg"]   # spilled U^/y^ go out once and come back (almost exactly) once


@pytest.mark.parametrize("slots", [8, 12, 48])
def test_program_op_small_circuits(oracle, golden, slots):
    for key in ["farid_n_najm", "test_case_1", "voltage_divider", "basic_001", "high_pass_filter", "low_pass_filter", "inverter"]:
        kc, oc = kcirc(golden, key), ocirc(oracle, golden, key)
        plan = kb.Plan(kc, api.MODE_DC, scratch_slots=slots)
        em = Emulator(plan, 1)
        X = em.run(np.zeros((0, 1)))
        assert_close(X[0][:, 0], oc.solve_op(), what=key)


def test_program_dc_sweep_sequential_warm_start(oracle, golden):
    """dc.rs:58,112: a nonlinear DC sweep is sequential, every point warm-started from its predecessor.  The DC lane program run on the
    sequential launch schedule reproduces the reference's iteration history (oracle warm_start=True) to 1e-9 — the lane-parallel
    cold start only agrees within the Newton-Raphson tolerance."""
    kc, oc = kcirc(golden, "diode_iv_curve"), ocirc(oracle, golden, "diode_iv_curve")
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("V1"))
    X, em, its = run_sequential_sweep(Emulator, plan, 0.0, 0.125, 17)
    warm = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=True)
    cold = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=False)
    assert warm.shape == (17, plan.n) and not em.status.any()
    assert_close(X[:, : plan.n, 0], warm, rtol=1e-9, what="sequential sweep vs the reference's warm start")
    assert np.max(np.abs(warm - cold)) > 1e-9            # the two iteration histories really differ: the test can tell them apart
    # the `step` pseudo-unknown (dc.rs:97) forces a second iteration at every warm-started point (|r - (r-1)| = 1 > reltol*r + 1e-6)
    assert (its[1:, 0] >= 2).all()


def test_program_dc_sweep_sequential_inverter_golden(oracle, golden):
    """circuits/inverter/dc/result.parquet through the sequential schedule (NMOS stamps nothing, but the circuit counts as nonlinear)."""
    g = golden["parquet"]["inverter_dc"]
    kc = kcirc(golden, "inverter")
    s = g["spec"]
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index(s["element"]))
    num = len(g["columns"]["step"])
    X, em, _ = run_sequential_sweep(Emulator, plan, s["start"], s["step_size"], num)
    assert not em.status.any()
    for name, col in g["columns"].items():
        if name != "step":
            np.testing.assert_allclose(X[:, kc.index_map[name], 0], col, rtol=1e-12, atol=1e-15, err_msg=name)


def test_program_nr_diode(oracle, golden):
    kc, oc = kcirc(golden, "diode_iv_curve"), ocirc(oracle, golden, "diode_iv_curve")
    plan = kb.Plan(kc, api.MODE_DC)
    em = Emulator(plan, 1)
    X = em.run(np.zeros((0, 1)))
    ref, it = oc.solve_op(return_iterations=True)
    assert_close(X[0][:, 0], ref, what="diode OP")
    assert em.iters[0] == it and em.status[0] == 0


def test_program_ac(oracle, golden):
    for key, n in [("low_pass_filter", 0), ("high_pass_filter", 0), ("rc", 60)]:
        text = netgen.rc_ladder(n) if key == "rc" else golden["netlists"][key]
        kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
        freqs = oracle.generate_frequencies("Decade", 3, 1.0, 1e7)
        plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e3, scratch_slots=14)
        em = Emulator(plan, len(freqs))
        X = em.run(freqs[None, :])
        _, ref = oc.solve_ac(freqs)
        assert_close(X[0].T, ref, what=key)


def test_program_transient_linear(oracle, golden):
    for key, h, stop in [("dual_rc_ladder", 50e-6, 5e-3), ("low_pass_filter_transient", 50e-6, 3e-3), ("high_pass_filter_transient", 10e-6, 1e-3)]:
        kc, oc = kcirc(golden, key), ocirc(oracle, golden, key)
        pdc = kb.Plan(kc, api.MODE_DC)
        ptr = kb.Plan(kc, api.MODE_TRAN, time_step=h)
        t, ref = oc.solve_tran(h, stop)
        rows = len(t)
        X = Emulator(pdc, 1, rows=rows).run(np.zeros((0, 1)), row_begin=0, row_end=1)
        em = Emulator(ptr, 1, rows=rows, h=h)
        assert em.split
        X = em.run(np.zeros((0, 1)), X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
        assert_close(X[:, :, 0], ref, what=key)


def test_program_transient_nonlinear(oracle, golden):
    kc, oc = kcirc(golden, "rectifier"), ocirc(oracle, golden, "rectifier")
    h, stop = 50e-6, 5e-3
    pdc = kb.Plan(kc, api.MODE_DC)
    ptr = kb.Plan(kc, api.MODE_TRAN, time_step=h)
    t, ref = oc.solve_tran(h, stop)
    rows = len(t)
    X = Emulator(pdc, 1, rows=rows).run(np.zeros((0, 1)), row_begin=0, row_end=1)
    em = Emulator(ptr, 1, rows=rows, h=h)
    assert not em.split
    X = em.run(np.zeros((0, 1)), X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
    assert_close(X[:, :, 0], ref, rtol=1e-7, what="rectifier")
    assert not em.status.any()


def test_program_monte_carlo_transient(oracle, golden):
    """BASELINE config 4 shape: dual_rc_ladder, R1 R2 C1 C2 varied per lane (splitmix64 rule of SURVEY §8d)."""
    text = golden["netlists"]["dual_rc_ladder"]
    kc = kb.parse_circuit_description(text)
    names = ["R1", "R2", "C1", "C2"]
    varied = [kc.element_index(nm) for nm in names]
    lanes, h, nsteps = 5, 5e-6, 40
    vals = np.array([[kc.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(lanes)] for j, e in enumerate(varied)])
    pdc = kb.Plan(kc, api.MODE_DC, varied=varied)
    ptr = kb.Plan(kc, api.MODE_TRAN, varied=varied, time_step=h)
    _, _, iov = input_layout(ptr)
    IN = np.zeros((ptr.info.n_inputs, lanes))
    for k, pl in enumerate(iov):
        IN[pl] = vals[k]
    rows = nsteps + 1
    X = Emulator(pdc, lanes, rows=rows).run(IN, row_begin=0, row_end=1)
    X = Emulator(ptr, lanes, rows=rows, h=h).run(IN, X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
    for l in range(lanes):
        oc = oracle.Circuit(text)
        for j, nm in enumerate(names):
            oc.set_element_value(oc.element_index(nm), vals[j, l])
        _, ref = oc.solve_tran(h, nsteps * h)
        assert_close(X[:, :, l], ref, what=f"lane {l}")


def test_program_mesh_with_fill(oracle):
    """A small RLC mesh: fill-in, structurally zero diagonals (L branches in DC), spills with a tiny scratch."""
    text = netgen.rlc_mesh(5, 6, transient=True)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(kc, api.MODE_DC, scratch_slots=10)
    assert plan.info.nnz_lu >= plan.info.nnz
    X = Emulator(plan, 1).run(np.zeros((0, 1)))
    assert_close(X[0][:, 0], oc.solve_op(), what="mesh OP")
    text = netgen.rlc_mesh(5, 6, transient=False)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = np.array([1.0, 1e3, 1e6, 1e9])
    plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e4, scratch_slots=12)
    X = Emulator(plan, len(freqs)).run(freqs[None, :])
    _, ref = oc.solve_ac(freqs)
    assert_close(X[0].T, ref, what="mesh AC")


# ---- chain executor tables ---------------------------------------------------------------------------------------
def test_chain_plan_selected_for_ladders(golden):
    from vm_emulator import run_chain  # noqa: F401
    ladder = kb.parse_circuit_description(netgen.resistor_ladder(300))
    assert kb.Plan(ladder, api.MODE_DC, sweep_element=0).info.executor == 1
    assert kb.Plan(ladder, api.MODE_DC, sweep_element=0, force_lane_vm=True).info.executor == 0
    assert kb.Plan(kb.parse_circuit_description(netgen.rc_ladder(300)), api.MODE_AC).info.executor == 1
    assert kb.Plan(kcirc(golden, "low_pass_filter"), api.MODE_AC).info.executor == 1          # n = 3 path
    assert kb.Plan(kcirc(golden, "farid_n_najm"), api.MODE_DC).info.executor == 0             # not a path graph
    assert kb.Plan(kcirc(golden, "diode_iv_curve"), api.MODE_DC).info.executor == 0           # nonlinear -> lane VM
    assert kb.Plan(kcirc(golden, "dual_rc_ladder"), api.MODE_TRAN, time_step=1e-6).info.executor == 0
    mesh = kb.parse_circuit_description(netgen.rlc_mesh(4, 4))
    assert kb.Plan(mesh, api.MODE_AC).info.executor == 0


def test_chain_terms_dc_sweep(oracle):
    from vm_emulator import run_chain
    kc, oc = kb.parse_circuit_description(netgen.resistor_ladder(500)), oracle.Circuit(netgen.resistor_ladder(500))
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=0)
    vals = np.linspace(-2.0, 3.0, 7)
    X, status = run_chain(plan, vals[None, :], False)
    assert_close(X.T, oc.solve_dc("V1", -2.0, 3.0, 5.0 / 6.0), what="chain DC sweep")
    assert not status.any()
    assert run_chain.last_fixed_rows >= 498          # all but the source rows are fixed-shape records
    assert plan.info.hbm_bytes_per_point == 8 * (5 * 501 + 2) + 8      # one row (the source's KCL row) owns a u2 plane


def test_chain_terms_ac_and_monte_carlo(oracle, golden):
    from vm_emulator import run_chain
    text = netgen.rc_ladder(120)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = oracle.generate_frequencies("Decade", 2, 1.0, 1e8)
    plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e4)
    X, status = run_chain(plan, freqs[None, :], True)
    _, ref = oc.solve_ac(freqs)
    assert_close(X.T, ref, what="chain AC")
    # an RLC line: the weighted matching pairs inductor-branch equations with node voltages (better pivots), which takes
    # it out of the (1,2)-band class -> general lane VM, same answers
    text = "V1 a 0 1 AC 1\nL1 a b 1u\nC1 b 0 1n\nL2 b c 2u\nC2 c 0 1n\nR1 c 0 50\n"
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e6)
    freqs = np.array([1.0, 1e5, 3e6, 1e9])
    if plan.info.executor == 1:
        X, _ = run_chain(plan, freqs[None, :], True)
    else:
        X = Emulator(plan, len(freqs)).run(freqs[None, :])[0]
    assert_close(X.T, oc.solve_ac(freqs)[1], what="RLC line")
    # Monte-Carlo: every resistor of a ladder varied + AC with varied capacitors
    text = netgen.rc_ladder(12)
    kc = kb.parse_circuit_description(text)
    varied = [kc.element_index(f"R{k}") for k in (1, 5, 12)] + [kc.element_index(f"C{k}") for k in (2, 7)]
    lanes = 4
    rng = np.random.default_rng(3)
    vals = np.array([[kc.element_value(e) * (0.9 + 0.2 * rng.random()) for _ in range(lanes)] for e in varied])
    plan = kb.Plan(kc, api.MODE_AC, varied=varied)
    assert plan.info.executor == 1
    _, fi, iov = input_layout(plan)
    IN = np.zeros((plan.info.n_inputs, lanes))
    f = np.array([10.0, 1e3, 1e5, 1e7])
    IN[fi] = f
    for k, pl in enumerate(iov):
        IN[pl] = vals[k]
    X, _ = run_chain(plan, IN, True)
    for l in range(lanes):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        assert_close(X[:, l], oc.solve_ac([f[l]])[1][0], what=f"chain MC lane {l}")


# ---- block-partitioned chain kernel tables (chain_block_kernel.cuh) ------------------------------------------------
def test_chain_block_tables_ladder_5000(oracle):
    """BASELINE config 2 shape: T = 256 threads x M = 20 rows; the numpy restatement of the kernel against the oracle."""
    from vm_emulator import run_chain_block, run_chain, chain_block_info
    text = netgen.resistor_ladder(5000)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=0)
    valid, T, M, n_dict, diff, why = chain_block_info(plan)
    assert valid and (T, M) == (256, 21), why
    assert n_dict <= 6 and diff <= 5e-10            # a ladder has a handful of distinct row records
    assert plan.info.hbm_bytes_per_point == 8 * 5001 + 8     # inputs in, n results out: SURVEY 8(d)'s fused floor
    vals = np.array([-1.5, 0.25, 3.0])
    X, status = run_chain_block(plan, vals[None, :], False)
    assert not status.any()
    assert_close(X.T, np.array([oc_solve_dc_point(oc, v) for v in vals]), what="block chain DC")
    Xs, _ = run_chain(plan, vals[None, :], False)
    assert_close(X, Xs, rtol=5e-10, what="block vs sequential chain")


def oc_solve_dc_point(oc, v):
    return oc.solve_dc("V1", v, v, 1.0)[0]


def test_chain_block_tables_rc_ladder_ac(oracle):
    """BASELINE config 3 shape (complex): T = 512 x M = 10, frequencies across eight decades."""
    from vm_emulator import run_chain_block, chain_block_info
    text = netgen.rc_ladder(5000)
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(kc, api.MODE_AC, nominal_frequency=1e4)
    valid, T, M, _, _, why = chain_block_info(plan)
    assert valid and (T, M) == (512, 10), why
    freqs = np.array([1.0, 1e3, 1e6, 1e8])
    X, status = run_chain_block(plan, freqs[None, :], True)
    assert not status.any()
    assert_close(X.T, oc.solve_ac(freqs)[1], what="block chain AC")


def test_chain_block_small_forced_monte_carlo_and_order(oracle, monkeypatch):
    """KRB_CHAIN_BLOCK=1 admits small systems (mostly padding rows): Monte-Carlo term runs on every row, a pivot order that is
    not the unknown order, and the default (n < 512 keeps the sequential kernel)."""
    from vm_emulator import run_chain_block, run_chain, chain_block_info
    text = netgen.rc_ladder(40)
    kc = kb.parse_circuit_description(text)
    assert not chain_block_info(kb.Plan(kc, api.MODE_AC))[0]
    monkeypatch.setenv("KRB_CHAIN_BLOCK", "1")
    varied = [kc.element_index(f"R{k}") for k in range(1, 41)] + [kc.element_index(f"C{k}") for k in (2, 7, 40)]
    lanes = 3
    rng = np.random.default_rng(5)
    vals = np.array([[kc.element_value(e) * (0.9 + 0.2 * rng.random()) for _ in range(lanes)] for e in varied])
    plan = kb.Plan(kc, api.MODE_AC, varied=varied)
    valid, T, M, _, _, why = chain_block_info(plan)
    assert valid and (T, M) == (512, 4), why
    _, fi, iov = input_layout(plan)
    IN = np.zeros((plan.info.n_inputs, lanes))
    f = np.array([10.0, 1e4, 1e7])
    IN[fi] = f
    for k, pl in enumerate(iov):
        IN[pl] = vals[k]
    X, status = run_chain_block(plan, IN, True)
    assert not status.any()
    for l in range(lanes):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        assert_close(X[:, l], oc.solve_ac([f[l]])[1][0], what=f"block chain MC lane {l}")
    # resistors listed from the grounded end, source last: unknown order is the reverse of the elimination order
    lines = [f"R{k} n{k} n{k + 1} 1k" for k in range(30, 0, -1)] + ["R31 n31 0 1k", "V1 n1 0 2"]
    text = "\n".join(lines) + "\n"
    kc, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=kc.element_index("V1"))
    if plan.info.executor == 1 and chain_block_info(plan)[0]:
        vals = np.array([0.5, 2.0])
        X, _ = run_chain_block(plan, vals[None, :], False)
        for l, v in enumerate(vals):
            assert_close(X[:, l], oc.solve_dc("V1", v, v, 1.0)[0], what="block chain, permuted order")
        Xs, _ = run_chain(plan, vals[None, :], False)
        assert_close(X, Xs, rtol=1e-11, what="block vs sequential, permuted order")
    # KRB_CHAIN_BLOCK=0 switches the tables off
    monkeypatch.setenv("KRB_CHAIN_BLOCK", "0")
    assert not chain_block_info(kb.Plan(kb.parse_circuit_description(netgen.resistor_ladder(5000)), api.MODE_DC, sweep_element=0))[0]


# ---- trapezoidal / lag-free integrators (north-star extension; not reference behaviour) ---------------------------
def test_trapezoidal_oracle_vs_closed_form(oracle):
    """The non-reference integrators of the oracle are validated against the closed-form RC response to a sine source:
    trapezoidal is second order (error drops ~4x per halving of h) and far more accurate than backward Euler."""
    text = "V1 in 0 SIN(0 1 100 0 0 0)\nR1 in out 1k\nC1 out 0 1u\n"
    oc = oracle.Circuit(text)
    w, rc_ = 2 * np.pi * 100.0, 1e-3
    def exact(t):
        a = w * rc_
        return (np.sin(w * t) - a * np.cos(w * t) + a * np.exp(-t / rc_)) / (1 + a * a)
    errs = {}
    for integ in (1, 2):
        for h in (100e-6, 50e-6, 25e-6):
            t, x = oc.solve_tran(h, 20e-3, integrator=integ)
            errs[(integ, h)] = np.max(np.abs(x[:, oc.index_map["V(out)"]] - exact(t)))
    assert errs[(1, 50e-6)] < 2e-4 and errs[(1, 50e-6)] < errs[(2, 50e-6)] / 20
    assert 3.5 < errs[(1, 100e-6)] / errs[(1, 50e-6)] < 4.5 and 3.5 < errs[(1, 50e-6)] / errs[(1, 25e-6)] < 4.5
    assert 1.7 < errs[(2, 100e-6)] / errs[(2, 50e-6)] < 2.3


@pytest.mark.parametrize("integ", [1, 2])
def test_program_transient_extensions(oracle, golden, integ):
    cases = [("dual_rc_ladder", 50e-6, 5e-3), ("high_pass_filter_transient", 10e-6, 1e-3), ("rectifier", 50e-6, 3e-3)]
    for key, h, stop in cases:
        kc, oc = kcirc(golden, key), ocirc(oracle, golden, key)
        pdc = kb.Plan(kc, api.MODE_DC)
        ptr = kb.Plan(kc, api.MODE_TRAN, time_step=h, integrator=integ)
        t, ref = oc.solve_tran(h, stop, integrator=integ)
        rows = len(t)
        em = Emulator(ptr, 1, rows=rows, h=h)
        X = np.zeros((rows, em.width, 1))
        X[:1, : kc.n] = Emulator(pdc, 1, rows=1).run(np.zeros((0, 1)))
        X = em.run(np.zeros((0, 1)), X=X, row_begin=1, guess_empty_at_iter0=False, pseudo_time=True)
        assert_close(X[:, : kc.n, 0], ref, rtol=1e-7 if key == "rectifier" else 1e-9, what=f"{key} integrator {integ}")


# ---- f3: subcircuit flattening (krets-parser/src/elements/subcircuit.rs:33-116, parser.rs:93-98) ---------------------
SUBCKT_FIXTURE = """* circuits/voltage_divider/subckt/voltage_divider.cir
xdiv1 10 7 0 vdivide
.subckt vdivide 1 2 3
r1 1 2 10K
r2 2 3 5K
.ends

V1 1 0 dc 10
.control
op
.endc
"""
SUBCKT_NESTED = """* crates/krets-parser/tests/test_parse_netlist.rs:78-102
xdiv1 10 7 0 vdivide

.subckt vdivide 1 2 3
xr1 1 2 myresistor
xr2 2 3 myresistor
.ends

.subckt myresistor plus minus
r0 plus minus 1k
.ends
"""


def _flatten_reference_way(text):
    """Python restatement of SubcircuitInstance::instantiate + map_sub_element for two-terminal elements (test oracle)."""
    defs, cur, top, control = {}, None, [], False
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line[0] in "%*":
            continue
        low = line.lower()
        if low.startswith(".control") or low.startswith(".endc"):
            control = low.startswith(".control")
            continue
        if control:
            continue
        if low.startswith(".subckt"):
            tok = line.split()
            cur = tok[1]
            defs[cur] = (tok[2:], [])
        elif low.startswith(".ends"):
            cur = None
        elif cur is not None:
            defs[cur][1].append(line.split())
        elif not low.startswith("."):
            top.append(line.split())

    def inst(name, nodes, dname):
        pins, elems = defs[dname]
        assert len(pins) == len(nodes)
        m = dict(zip(pins, nodes))
        out = []
        for tok in elems:
            if tok[0][0] in "xX":
                out += inst(name + "_" + tok[0][1:], [m.get(t, name + "_" + t) for t in tok[1:-1]], tok[-1])
            else:
                out.append((tok[0][0].upper() + name + "_" + tok[0][1:], m.get(tok[1], name + "_" + tok[1]), m.get(tok[2], name + "_" + tok[2])))
        return out

    flat = []
    for tok in top:
        if tok[0][0] in "xX":
            flat += inst(tok[0][1:], tok[1:-1], tok[-1])
        else:
            flat.append((tok[0][0].upper() + tok[0][1:], tok[1], tok[2]))
    return flat


@pytest.mark.parametrize("text,count", [(SUBCKT_FIXTURE, 3), (SUBCKT_NESTED, 2)])
def test_subckt_flattening_matches_reference_rules(oracle, text, count):
    circuit = kb.parse_circuit_description(text)
    oc = oracle.Circuit(text)                      # the C oracle restates subcircuit.rs independently of the host front end
    assert circuit.names == oc.names and circuit.num_elements == oc.num_elements
    for e in range(circuit.num_elements):
        el = circuit.element(e)
        assert (el.ip, el.im, -1 if el.kind == api.KIND_NMOS else el.ic) == oc.element_indices(e) and circuit.identifiers[e] == oc.element_ident(e)
    L = _lib.lib()
    idents = [L.krh_circuit_element_identifier(circuit._h, e).decode() for e in range(circuit.num_elements)]
    expect = _flatten_reference_way(text)
    assert circuit.num_elements == count == len(expect)
    assert idents == [e[0] for e in expect]
    # index map: parser.rs:105-125 applied to the flattened element list
    names, seen = [], set()
    for ident, plus, minus in expect:
        if ident[0] in "VLI":
            names.append(f"I({ident})")
        for node in (plus, minus):
            if node not in seen:
                seen.add(node)
                if node != "0":
                    names.append(f"V({node})")
    assert circuit.names == names
    for e, (ident, plus, minus) in enumerate(expect):
        el = circuit.element(e)
        assert el.ip == circuit.index_map.get(f"V({plus})", -1) and el.im == circuit.index_map.get(f"V({minus})", -1)


def test_subckt_internal_ground_is_prefixed_and_errors():
    # a "0" inside a definition that is not a pin becomes "<instance>_0" (map_sub_element, subcircuit.rs:96-103)
    c = kb.parse_circuit_description("V1 in 0 1\nxa in out cell\n.subckt cell a b\nr1 a b 1k\nr2 b 0 2k\n.ends\n")
    assert c.names == ["I(V1)", "V(in)", "V(out)", "V(a_0)"]
    for bad in ("xa 1 2 nodef\nR1 1 0 1\n", "xa 1 cell\n.subckt cell a b\nr1 a b 1\n.ends\n", "xa 1 2 cell\n.subckt cell a b\nq1 a b c\n.ends\n"):
        with pytest.raises(kb.KretsError):
            kb.parse_circuit_description(bad)
