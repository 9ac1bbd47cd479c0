"""GPU tier (-m gpu): parity of the CUDA path against the CPU oracle, through the C ABI.

Tolerance (north star): max relative error 1e-9 in f64; complex results are compared through |a-b|/|b| per unknown,
which bounds both the magnitude and the phase error.  The reference's own test assertions (abs 1e-3 / 1e-4) and the
full-precision golden parquets are replayed through the Solver-shaped API so these read like crates/krets-solver/tests.
"""
import math
import os

import numpy as np
import pytest

import netgen
from helpers import assert_close, rel_err

import krets_b200 as kb
from krets_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert kb.device_count() > 0, "GPU tier needs a CUDA device (the library has no CPU fallback)"


def both(oracle, golden, key):
    if key.startswith("resistor_ladder_"):
        text = netgen.resistor_ladder(int(key.rsplit("_", 1)[1]))
    else:
        text = golden["netlists"][key]
    return kb.parse_circuit_description(text), oracle.Circuit(text)


# ---- crates/krets-solver/tests/test_op.rs --------------------------------------------------------------------
@pytest.mark.parametrize("key", ["farid_n_najm", "test_case_1", "voltage_divider", "low_pass_filter", "high_pass_filter",
                                 "basic_001", "resistor_ladder_500", "diode_iv_curve"])
def test_op(oracle, golden, key):
    circuit, oc = both(oracle, golden, key)
    solver = kb.Solver(circuit, kb.SolverConfig.default())
    solution = solver.solve(kb.Analysis.Op()).into_op()
    for name, exp, tol, where in golden["assertions"]["op"][key]:
        assert abs(solution[name] - exp) < tol, (name, where)
    ref = oc.solve_op()
    assert_close([solution[nm] for nm in circuit.names], ref, what=key)


def test_op_resistor_ladder_5000(oracle):
    circuit, oc = kb.parse_circuit_description(netgen.resistor_ladder(5000)), oracle.Circuit(netgen.resistor_ladder(5000))
    solution = kb.Solver(circuit).solve(kb.Analysis.Op())
    assert_close(solution.array[0], oc.solve_op(), what="ladder 5000 OP")


# ---- tests/test_dc.rs + inverter golden ----------------------------------------------------------------------
def test_dc_voltage_divider(oracle, golden):
    a = golden["assertions"]["dc"]["voltage_divider"]
    circuit, oc = both(oracle, golden, "voltage_divider")
    solver = kb.Solver(circuit)
    solution = solver.solve(kb.Analysis.Dc(kb.DcAnalysis(a["element"], a["start"], a["stop"], a["step_size"]))).into_dc()
    assert len(solution) == a["len"]
    for r, checks in a["rows"].items():
        for name, exp, tol, where in checks:
            assert abs(solution[int(r)][name] - exp) < tol, (name, where)
    assert [row["step"] for row in solution] == [0.0, 1.0]
    # a second solve on the same Solver sees the original source value again (dc.rs:117-121)
    assert abs(solver.solve(kb.Analysis.Op()).into_op()["V(in)"] - 1.0) < 1e-12


def test_dc_errors(golden, oracle):
    circuit, _ = both(oracle, golden, "voltage_divider")
    solver = kb.Solver(circuit)
    with pytest.raises(kb.KretsError) as e:
        solver.solve(kb.Analysis.Dc(kb.DcAnalysis("V9", 0.0, 1.0, 1.0)))
    assert e.value.kind == "ElementNotFound"
    with pytest.raises(kb.KretsError) as e:
        solver.solve(kb.Analysis.Dc(kb.DcAnalysis("R1", 0.0, 1.0, 1.0)))
    assert e.value.kind == "InvalidElementFormat"


def test_dc_inverter_golden(oracle, golden):
    """circuits/inverter/dc/result.parquet (101 rows): NMOS no-op, NR loop with the `step` pseudo-unknown."""
    g = golden["parquet"]["inverter_dc"]
    circuit, _ = both(oracle, golden, "inverter")
    s = g["spec"]
    res = kb.Solver(circuit).solve(kb.Analysis.Dc(kb.DcAnalysis(s["element"], s["start"], s["stop"], s["step_size"])))
    assert res.array.shape[0] == 101 and res.axis.tolist() == g["columns"]["step"]
    for name, col in g["columns"].items():
        if name != "step":
            np.testing.assert_allclose(res.array[:, circuit.index_map[name]], col, rtol=1e-12, atol=1e-15, err_msg=name)


def test_dc_diode_sweep_reference_warm_start(oracle, golden):
    """Nonlinear DC sweep through the Solver: the reference's sequential sweep, every point warm-started from its predecessor
    (dc.rs:58,112) — parity with the oracle's warm-started sweep at 1e-9 (the iteration history is the reference's)."""
    circuit, oc = both(oracle, golden, "diode_iv_curve")
    res = kb.Solver(circuit).solve(kb.Analysis.Dc(kb.DcAnalysis("V1", 0.0, 2.0, 0.125)))
    warm = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=True)
    cold = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=False)
    assert res.array.shape == warm.shape and res.axis.tolist() == [float(i) for i in range(17)]
    assert_close(res.array, warm, rtol=1e-9, what="diode sweep vs the reference's warm-started sweep")
    assert np.max(np.abs(warm - cold)) > 1e-9     # the two iteration histories differ measurably: 1e-9 tells them apart


def test_dc_diode_sweep_parallel_cold_start(oracle, golden):
    """Extension (DcAnalysis.mode = DC_PARALLEL_COLD_START): the points of a nonlinear sweep as parallel lanes, every lane from the empty
    guess; parity is against the oracle run the same way, and within NR tolerance against the reference's sequential warm start."""
    circuit, oc = both(oracle, golden, "diode_iv_curve")
    res = kb.Solver(circuit).solve(kb.Analysis.Dc(kb.DcAnalysis("V1", 0.0, 2.0, 0.125, mode=api.DC_PARALLEL_COLD_START)))
    cold = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=False)
    warm = oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=True)
    assert_close(res.array, cold, rtol=1e-8, what="diode sweep vs cold-start oracle")
    assert np.max(np.abs(res.array - warm)) < 2e-3


def test_dc_sweep_sequential_monte_carlo_lanes(oracle, golden):
    """krb_solve_dc_sweep_sequential: Monte-Carlo lanes (R1 per lane), each a warm-started sequential sweep; interpreter and
    circuit-specialised kernel agree bit for bit; every lane equals the oracle's warm-started sweep of that variant."""
    circuit, oc = both(oracle, golden, "diode_iv_curve")
    r1 = circuit.element_index("R1")
    vals = np.array([[1000.0, 900.0, 1500.0, 330.0, 4700.0]])
    outs = []
    for jit in ("1", "0"):
        os.environ["KRB_VM_JIT"] = jit
        try:
            plan = kb.Plan(circuit, api.MODE_DC, sweep_element=circuit.element_index("V1"), varied=[r1])
            x, status, iters = plan.solve_dc_sweep_sequential(0.0, 0.125, 17, lanes=5, varied_values=vals, return_iterations=True)
        finally:
            os.environ.pop("KRB_VM_JIT", None)
        assert x.shape == (5, 17, circuit.n) and not status.any() and (iters >= 2).all()
        outs.append(x)
    assert np.array_equal(outs[0], outs[1])
    e = oc.element_index("R1")
    for l in range(5):
        oc.set_element_value(e, float(vals[0, l]))
        assert_close(outs[0][l], oc.solve_dc("V1", 0.0, 2.0, 0.125, warm_start=True), rtol=1e-9, what=f"lane {l}")
    oc.set_element_value(e, 1000.0)


def test_dc_sweep_sequential_rejects_plans_off_the_lane_vm(oracle):
    """Linear chain plans cannot depend on a warm start: the sequential entry point refuses them instead of running something else."""
    kc = kb.parse_circuit_description(netgen.resistor_ladder(600))
    plan = kb.Plan(kc, api.MODE_DC, sweep_element=0)
    with pytest.raises(kb.KretsError) as e:
        plan.solve_dc_sweep_sequential(0.0, 1.0, 3)
    assert e.value.kind == "Unsupported"


# ---- tests/test_ac.rs + golden parquet -------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["low_pass_filter", "high_pass_filter"])
def test_ac_single_point(oracle, golden, key):
    circuit, oc = both(oracle, golden, key)
    ac = kb.AcAnalysis(kb.AcSweep.Linear(1), 1000.0, 1000.0)
    solution = kb.Solver(circuit).solve(kb.Analysis.Ac(ac)).into_ac()[0]
    for name, re, im, tol, where in golden["assertions"]["ac"][key]:
        assert abs(solution[name].real - re) < tol and abs(solution[name].imag - im) < tol, (name, where)


def test_ac_sweep_closed_form_and_oracle(oracle, golden):
    circuit, oc = both(oracle, golden, "low_pass_filter")
    ac = kb.AcAnalysis(kb.AcSweep.Linear(100), 1.0, 1000.0)     # BASELINE config 1
    res = kb.Solver(circuit).solve(kb.Analysis.Ac(ac))
    assert res.array.shape == (100, 3)
    for row in res.into_ac():
        w = 2 * math.pi * row["frequency"].real * 1000.0 * 1e-6
        assert abs(row["V(out)"].real - 1 / (1 + w * w)) < 1e-3 and abs(row["V(out)"].imag + w / (1 + w * w)) < 1e-3
    fr, ref = oc.solve_ac(oracle.generate_frequencies("Linear", 100, 1.0, 1000.0))
    assert res.axis.tobytes() == fr.tobytes()
    assert_close(res.array, ref, what="LP AC 100 pts")


def test_ac_golden_parquet(oracle, golden):
    g = golden["parquet"]["low_pass_filter_ac"]
    circuit, _ = both(oracle, golden, "low_pass_filter_ac")
    s = g["spec"]
    res = kb.Solver(circuit).solve(kb.Analysis.Ac(kb.AcAnalysis(kb.AcSweep(s["sweep"], s["total_points"]), s["fstart"], s["fstop"])))
    cols = g["columns"]
    assert res.axis.tolist() == cols["frequency"]
    for name in ("I(V1)", "V(in)", "V(out)"):
        v = res.array[:, circuit.index_map[name]]
        np.testing.assert_allclose(np.abs(v), cols[name + "_mag"], rtol=1e-12, err_msg=name)
        np.testing.assert_allclose(np.degrees(np.arctan2(v.imag, v.real)), cols[name + "_phase_deg"], rtol=1e-11, atol=1e-11, err_msg=name)


def test_ac_with_current_source_fails(oracle, golden):
    circuit, _ = both(oracle, golden, "basic_001")
    with pytest.raises(kb.KretsError) as e:
        kb.Solver(circuit).solve(kb.Analysis.Ac(kb.AcAnalysis(kb.AcSweep.Linear(1), 1000.0, 1000.0)))
    assert e.value.kind == "DecompositionFailed"


def test_ac_diode_small_signal(oracle, golden):
    circuit, oc = both(oracle, golden, "diode_iv_curve")
    res = kb.Solver(circuit).solve(kb.Analysis.Ac(kb.AcAnalysis(kb.AcSweep.Decade(5), 10.0, 1e5)))
    fr, ref = oc.solve_ac(oracle.generate_frequencies("Decade", 5, 10.0, 1e5))
    # V1 has no AC amplitude -> all-zero response; compare absolutely
    assert np.max(np.abs(res.array - ref)) < 1e-15


# ---- tests/test_transient.rs ----------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["dual_rc_ladder", "low_pass_filter_transient", "high_pass_filter_transient"])
def test_transient_linear(oracle, golden, key):
    a = golden["assertions"]["transient"][key]
    circuit, oc = both(oracle, golden, key)
    res = kb.Solver(circuit).solve(kb.Analysis.Transient(kb.TransientAnalysis(a["time_step"], a["stop_time"])))
    rows = res.into_transient()
    for r, checks in a["rows"].items():
        for name, exp, tol, where in checks:
            assert abs(rows[int(r)][name] - exp) < tol, (name, where)
    t, ref = oc.solve_tran(a["time_step"], a["stop_time"])
    assert res.axis.tobytes() == t.tobytes()
    assert_close(res.array, ref, what=key)


def test_transient_rectifier(oracle, golden):
    circuit, oc = both(oracle, golden, "rectifier")
    res = kb.Solver(circuit).solve(kb.Analysis.Transient(kb.TransientAnalysis(50e-6, 50e-3)))
    t, ref = oc.solve_tran(50e-6, 50e-3)
    assert res.array.shape == ref.shape == (1001, circuit.n)
    assert_close(res.array, ref, rtol=1e-6, what="rectifier (NR per step; exp() differs in the last ulp)")


def test_max_iterations_error(oracle, golden):
    circuit, oc = both(oracle, golden, "diode_iv_curve")
    cfg = kb.SolverConfig(maximum_iterations=2)
    with pytest.raises(kb.KretsError) as e:
        kb.Solver(circuit, cfg).solve(kb.Analysis.Op())
    assert e.value.kind == "MaximumIterationsExceeded"
    with pytest.raises(oracle.OracleError) as e2:
        oc.solve_op(oracle.SolverConfig(maximum_iterations=2))
    assert e2.value.kind == "MaximumIterationsExceeded"


# ---- lane batches through krb_solve_* ---------------------------------------------------------------------------
@pytest.mark.parametrize("lanes", [1, 63, 64, 1000])
def test_dc_sweep_lanes_host_arrays(oracle, lanes):
    text = netgen.resistor_ladder(500)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0)
    step = 2.0 ** -6
    vals = np.arange(lanes) * step
    out, status = plan.solve_dc_sweep(vals)
    assert not status.any()
    ref = oc.solve_dc("V1", 0.0, (lanes - 1) * step, step, nthreads=8) if lanes > 1 else oc.solve_dc("V1", 0.0, 0.0, step)
    assert_close(out, ref, what=f"ladder DC sweep {lanes} lanes")
    out2, _ = plan.solve_dc_sweep(vals, layout=api.LAYOUT_LANE_INTERLEAVED, out=np.zeros((circuit.n, lanes)))
    assert np.array_equal(out2.T, out)


def test_empty_batch():
    circuit = kb.parse_circuit_description(netgen.resistor_ladder(10))
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0)
    out, status = plan.solve_dc_sweep(np.zeros(0))
    assert out.shape == (0, 11)


def test_dc_sweep_device_tensors_and_small_scratch(oracle):
    import torch
    text = netgen.resistor_ladder(500)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    lanes = 4097
    vals = torch.linspace(-1.0, 1.0, lanes, dtype=torch.float64, device="cuda")
    ref1 = oc.solve_op()
    for slots in (0, 8):
        plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0, scratch_slots=slots)
        out, status = plan.solve_dc_sweep(vals)
        torch.cuda.synchronize()
        assert not status.any().item()
        # linearity: x(lane) = v_lane * x(1 V) (V1 nominal is 1 V)
        exp = vals.cpu().numpy()[:, None] * ref1[None, :]
        assert_close(out.cpu().numpy(), exp, what=f"linearity, scratch={slots}")


def test_ac_rc_ladder_lanes(oracle):
    text = netgen.rc_ladder(300)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = kb.generate_frequencies("Decade", 250, 1.0, 1e8)
    assert len(freqs) == 2001
    plan = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e4)
    out, status = plan.solve_ac(freqs)
    assert not status.any()
    _, ref = oc.solve_ac(freqs, nthreads=8)
    assert_close(out, ref, what="RC ladder AC 2001 lanes")


def test_monte_carlo_transient(oracle, golden):
    """BASELINE config 4 at reduced size: dual_rc_ladder, R1 R2 C1 C2 varied, 64 lanes x 400 steps of 5 us."""
    text = golden["netlists"]["dual_rc_ladder"]
    circuit = kb.parse_circuit_description(text)
    names = ["R1", "R2", "C1", "C2"]
    varied = [circuit.element_index(nm) for nm in names]
    lanes, h, nsteps = 64, 5e-6, 400
    vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(lanes)] for j, e in enumerate(varied)])
    pdc = kb.Plan(circuit, api.MODE_DC, varied=varied)
    ptr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h)
    out, status = ptr.solve_transient(pdc, nsteps, lanes, varied_values=np.ascontiguousarray(vals))
    assert not status.any() and out.shape == (lanes, nsteps + 1, circuit.n)
    for l in (0, 1, 17, 63):
        oc = oracle.Circuit(text)
        for j, nm in enumerate(names):
            oc.set_element_value(oc.element_index(nm), vals[j, l])
        _, ref = oc.solve_tran(h, nsteps * h)
        assert_close(out[l], ref, what=f"MC transient lane {l}")


def test_monte_carlo_op_resistors(oracle):
    text = netgen.resistor_ladder(40)
    circuit = kb.parse_circuit_description(text)
    varied = list(range(1, 41))
    lanes = 33
    rng = np.random.default_rng(7)
    vals = 1000.0 * (0.95 + 0.1 * rng.random((40, lanes)))
    plan = kb.Plan(circuit, api.MODE_DC, varied=varied)
    out, status = plan.solve_op(lanes, varied_values=vals)
    assert not status.any()
    for l in (0, 5, 32):
        oc = oracle.Circuit(text)
        for j, e in enumerate(varied):
            oc.set_element_value(e, vals[j, l])
        assert_close(out[l], oc.solve_op(), what=f"MC OP lane {l}")


def test_zero_pivot_lane_status():
    """A lane whose resistor value is +inf gets g = 0 on an isolated node -> exactly zero pivot -> per-lane status, other lanes fine."""
    circuit = kb.parse_circuit_description("V1 a 0 1\nR1 a b 1k\nR2 b 0 1k\n")
    plan = kb.Plan(circuit, api.MODE_DC, varied=[1, 2])
    vals = np.array([[1000.0, np.inf], [1000.0, np.inf]])
    out, status = plan.solve_op(2, varied_values=vals)
    assert status[0] == 0 and abs(out[0][circuit.index_map["V(b)"]] - 0.5) < 1e-12
    assert status[1] & api.LANE_ZERO_PIVOT


def test_mesh_with_fill(oracle):
    text = netgen.rlc_mesh(8, 9, transient=True)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    assert_close(kb.Solver(circuit).solve(kb.Analysis.Op()).array[0], oc.solve_op(), what="mesh OP")
    res = kb.Solver(circuit).solve(kb.Analysis.Transient(kb.TransientAnalysis(1e-7, 2e-5)))
    _, ref = oc.solve_tran(1e-7, 2e-5)
    assert_close(res.array, ref, rtol=1e-8, what="mesh transient")
    text = netgen.rlc_mesh(8, 9, transient=False)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    res = kb.Solver(circuit).solve(kb.Analysis.Ac(kb.AcAnalysis(kb.AcSweep.Decade(4), 1.0, 1e9)))
    _, ref = oc.solve_ac(oracle.generate_frequencies("Decade", 4, 1.0, 1e9))
    assert_close(res.array, ref, what="mesh AC")


# ---- BASELINE sizes: size-independent properties ---------------------------------------------------------------
def test_config2_full_size_linearity(oracle):
    """resistor_ladder_5000, 65 536 sweep values: x(lane) must equal v_lane * x(1 V); one oracle solve anchors it."""
    import torch
    text = netgen.resistor_ladder(5000)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0)
    lanes = 65536
    vals = torch.arange(lanes, dtype=torch.float64, device="cuda") * 2.0 ** -16
    out, status = plan.solve_dc_sweep(vals)
    torch.cuda.synchronize()
    assert int(status.abs().max().item()) == 0
    x1 = torch.from_numpy(oc.solve_op()).cuda()
    exp = vals[:, None] * x1[None, :]
    scale = torch.maximum(exp.abs(), 1e-6 * exp.abs().amax(dim=1, keepdim=True))
    scale[0] = 1.0                                # lane 0 is the all-zero solution
    err = ((out - exp).abs() / scale).max().item()
    assert err <= 1e-9, err
    assert out[0].abs().max().item() == 0.0


def test_config3_chunk_rc_ladder_5000(oracle):
    """RC ladder 5000 (BASELINE config 3), an 8192-lane slice spread over the whole 1e6-point decade grid: EVERY lane and every
    unknown against the oracle (the oracle needs ~10-20 s on the box's host threads).
    helpers.rel_err measures unknowns more than six orders below the lane's largest one against that floor: at high frequency the
    ladder attenuates by hundreds of decades, so most of the 5001 unknowns of such a lane are compared absolutely (they underflow
    towards zero in both solvers)."""
    import os
    import torch
    text = netgen.rc_ladder(5000)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    allf = kb.generate_frequencies("Decade", 125000, 1.0, 1e8)
    assert abs(len(allf) - 1000001) <= 1
    sel = allf[:: len(allf) // 8192][:8192].copy()
    plan = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e4)
    f_dev = torch.from_numpy(sel).cuda()
    out, status = plan.solve_ac(f_dev)
    torch.cuda.synchronize()
    assert int(status.abs().max().item()) == 0
    host = out.cpu().numpy()
    _, ref = oc.solve_ac(sel, nthreads=os.cpu_count() or 1)
    assert ref.shape == host.shape
    worst = max(rel_err(host[l], ref[l]) for l in range(len(sel)))      # per lane: each lane has its own scale
    assert worst <= 1e-9, f"RC ladder 5000, all {len(sel)} lanes: max relative error {worst:.3e}"
    # the part of each lane that carries signal, compared purely relatively (no floor): unknowns within 1e-6 of the lane's maximum
    big = np.abs(ref) >= 1e-6 * np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(host - ref)[big] / np.abs(ref)[big]) <= 1e-9
    # residual of the source row: V(1) = 1 for every lane (row I(V1): V(1) = ac amplitude)
    assert np.max(np.abs(host[:, circuit.index_map["V(1)"]] - 1.0)) < 1e-12


# ---- both executors on the same inputs ---------------------------------------------------------------------------
def test_executors_agree_on_ladder(oracle):
    """chain kernel (default for ladders) and the general lane VM (forced) against the oracle and each other."""
    text = netgen.resistor_ladder(500)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    vals = np.linspace(-3.0, 3.0, 257)
    ref = oc.solve_dc("V1", -3.0, 3.0, 6.0 / 256, nthreads=8)
    outs = []
    for force in (False, True):
        plan = kb.Plan(circuit, api.MODE_DC, sweep_element=0, force_lane_vm=force)
        assert plan.info.executor == (0 if force else 1)
        out, status = plan.solve_dc_sweep(vals)
        assert not status.any()
        assert_close(out, ref, what=f"ladder DC force_vm={force}")
        outs.append(out)
    assert_close(outs[0], outs[1], rtol=1e-12, what="chain vs lane VM")
    text = netgen.rc_ladder(200)
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    freqs = kb.generate_frequencies("Decade", 20, 1.0, 1e8)
    _, ref = oc.solve_ac(freqs, nthreads=8)
    for force in (False, True):
        out, status = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1e4, force_lane_vm=force).solve_ac(freqs)
        assert not status.any()
        assert_close(out, ref, what=f"RC ladder AC force_vm={force}")


def test_chain_monte_carlo_ac(oracle):
    text = netgen.rc_ladder(30)
    circuit = kb.parse_circuit_description(text)
    varied = [circuit.element_index(f"R{k}") for k in range(1, 31)] + [circuit.element_index(f"C{k}") for k in range(1, 31)]
    lanes = 70
    rng = np.random.default_rng(11)
    vals = np.array([[circuit.element_value(e) * (0.95 + 0.1 * rng.random()) for _ in range(lanes)] for e in varied])
    freqs = np.logspace(1, 7, lanes)
    plan = kb.Plan(circuit, api.MODE_AC, varied=varied)
    assert plan.info.executor == 1
    out, status = plan.solve_ac(freqs, varied_values=np.ascontiguousarray(vals))
    assert not status.any()
    for l in (0, 33, 69):
        oc = oracle.Circuit(text)
        for k, e in enumerate(varied):
            oc.set_element_value(e, vals[k, l])
        assert_close(out[l], oc.solve_ac([freqs[l]])[1][0], what=f"chain MC AC lane {l}")


# ---- trapezoidal / lag-free backward Euler (extension; checked against the oracle's restatement) -----------------
@pytest.mark.parametrize("integ", [1, 2])
def test_transient_extension_integrators(oracle, golden, integ):
    for key, h, stop, tol in [("dual_rc_ladder", 50e-6, 20e-3, 1e-9), ("low_pass_filter_transient", 50e-6, 5e-3, 1e-9),
                              ("high_pass_filter_transient", 10e-6, 2e-3, 1e-9), ("rectifier", 50e-6, 20e-3, 1e-6)]:
        circuit, oc = both(oracle, golden, key)
        res = kb.Solver(circuit).solve(kb.Analysis.Transient(kb.TransientAnalysis(h, stop, integrator=integ)))
        t, ref = oc.solve_tran(h, stop, integrator=integ)
        assert res.array.shape == ref.shape
        assert_close(res.array, ref, rtol=tol, what=f"{key} integrator {integ}")


def test_monte_carlo_transient_trapezoidal(oracle, golden):
    """BASELINE config 4 as written (trapezoidal steps x Monte-Carlo R/C variants) at reduced size."""
    text = golden["netlists"]["dual_rc_ladder"]
    circuit = kb.parse_circuit_description(text)
    names = ["R1", "R2", "C1", "C2"]
    varied = [circuit.element_index(nm) for nm in names]
    lanes, h, nsteps = 96, 5e-6, 300
    vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(lanes)] for j, e in enumerate(varied)])
    pdc = kb.Plan(circuit, api.MODE_DC, varied=varied)
    ptr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h, integrator=api.INTEGRATOR_TRAP)
    out, status = ptr.solve_transient(pdc, nsteps, lanes, varied_values=np.ascontiguousarray(vals))
    assert not status.any() and out.shape == (lanes, nsteps + 1, circuit.n)
    out_li, _ = ptr.solve_transient(pdc, nsteps, lanes, varied_values=np.ascontiguousarray(vals), layout=api.LAYOUT_LANE_INTERLEAVED)
    assert np.array_equal(np.transpose(out_li, (2, 0, 1)), out)
    for l in (0, 41, 95):
        oc = oracle.Circuit(text)
        for j, nm in enumerate(names):
            oc.set_element_value(oc.element_index(nm), vals[j, l])
        _, ref = oc.solve_tran(h, nsteps * h, integrator=1)
        assert_close(out[l], ref, what=f"MC trapezoidal lane {l}")


# ---- f3: subcircuit instances are flattened on the host exactly as the reference does (subcircuit.rs:33-116) -------------
def test_op_with_subcircuits(oracle):
    text = ("xdiv1 10 7 0 vdivide\n.subckt vdivide 1 2 3\nr1 1 2 10K\nr2 2 3 5K\n.ends\nV1 10 0 dc 10\n"
            "xtop 7 out stage\n.subckt stage a b\nxr a mid res\nxs mid b res\nr9 b 0 4k\n.ends\n.subckt res p m\nr0 p m 1k\n.ends\n")
    flat = ("Rdiv1_1 10 7 10K\nRdiv1_2 7 0 5K\nV1 10 0 dc 10\nRtop_r_0 7 top_mid 1k\nRtop_s_0 top_mid out 1k\nRtop_9 out top_0 4k\n")
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(flat)
    assert circuit.names == oc.names == oracle.Circuit(text).names          # the oracle flattens subcircuits too
    assert_close(oracle.Circuit(text).solve_op(), oc.solve_op(), what="oracle: flattened by the parser vs by hand")
    sol = kb.Solver(circuit).solve(kb.Analysis.Op())
    # "top_0": a ground inside a definition that is not a pin is an ordinary internal node (reference behaviour), so the
    # second stage floats at V(7) — the oracle on the hand-flattened netlist agrees
    assert_close(sol.array[0], oc.solve_op(), what="subcircuit OP")
    x = dict(zip(circuit.names, sol.array[0]))
    assert abs(x["V(7)"] - 10.0 * 5.0 / 15.0) < 1e-12 and abs(x["V(top_0)"] - x["V(7)"]) < 1e-12


# ---- dense transient executor (n <= 8, linear): native stepping kernel vs the lane VM vs the oracle ------------------------
@pytest.mark.parametrize("integ", [api.INTEGRATOR_BE_KRETS, api.INTEGRATOR_TRAP, api.INTEGRATOR_BE])
def test_dense_transient_kernel_matches_lane_vm_and_oracle(oracle, golden, integ, monkeypatch):
    import torch
    circuit, oc = both(oracle, golden, "dual_rc_ladder")
    names = ["R1", "R2", "C1", "C2"]
    varied = [circuit.element_index(nm) for nm in names]
    lanes, h, steps = 70, 5e-6, 400
    vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(lanes)] for j, e in enumerate(varied)])
    pdc = kb.Plan(circuit, api.MODE_DC, varied=varied)
    ptr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h, integrator=integ)
    out_dense, st = ptr.solve_transient(pdc, steps, lanes, varied_values=vals)                       # host buffers
    assert not st.any()
    dv = torch.from_numpy(vals).cuda()
    out_dev, st_dev = ptr.solve_transient(pdc, steps, lanes, varied_values=dv)                       # device buffers, direct delivery
    assert not st_dev.any().item() and np.array_equal(out_dev.cpu().numpy(), out_dense)
    out_li, _ = ptr.solve_transient(pdc, steps, lanes, varied_values=vals, layout=api.LAYOUT_LANE_INTERLEAVED)
    assert np.array_equal(np.transpose(out_li, (2, 0, 1)), out_dense)
    monkeypatch.setenv("KRB_DENSE_TRAN_LU", "1")           # forward / backward substitution every step instead of the propagator
    out_lu, st_lu = ptr.solve_transient(pdc, steps, lanes, varied_values=vals)
    monkeypatch.delenv("KRB_DENSE_TRAN_LU")
    assert not st_lu.any()
    assert_close(out_dense, out_lu, rtol=1e-10, what="dense kernel: propagator steps vs LU steps")
    monkeypatch.setenv("KRB_NO_DENSE_TRAN", "1")
    pdc2 = kb.Plan(circuit, api.MODE_DC, varied=varied)
    ptr2 = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h, integrator=integ)
    out_vm, st2 = ptr2.solve_transient(pdc2, steps, lanes, varied_values=vals)
    assert not st2.any()
    assert_close(out_dense, out_vm, rtol=1e-10, what="dense kernel vs lane VM")
    for lane in (0, 1, lanes - 1):
        for j, e in enumerate(varied):
            oc.set_element_value(e, vals[j, lane])
        _, ref = oc.solve_tran(h, steps * h, integrator=integ)
        assert_close(out_dense[lane], ref, rtol=1e-9, what=f"dense transient integrator {integ} lane {lane}")


def test_config4_full_size_properties(golden):
    """BASELINE config 4 at full size (10^4 steps x 16 384 Monte-Carlo lanes): lane 0 is the nominal circuit; every lane is a
    passive RC ladder driven by a 0 -> 1 V step, so V(out) rises monotonically (backward Euler is monotone here), stays below
    V(in) = 1 V and is within 2 % of it after 50 ms (tau ~ 10 ms +- 10 %); the result is finite everywhere."""
    import torch
    circuit = kb.parse_circuit_description(golden["netlists"]["dual_rc_ladder"])
    varied = [circuit.element_index(nm) for nm in ["R1", "R2", "C1", "C2"]]
    lanes, steps = 16384, 10000
    vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(lanes)] for j, e in enumerate(varied)])
    pdc = kb.Plan(circuit, api.MODE_DC, varied=varied)
    ptr = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=5e-6)
    out, st = ptr.solve_transient(pdc, steps, lanes, varied_values=torch.from_numpy(vals).cuda())
    assert not st.any().item() and bool(torch.isfinite(out).all().item())
    nominal, _ = ptr.solve_transient(pdc, steps, 1, varied_values=torch.from_numpy(vals[:, :1].copy()).cuda())
    assert torch.equal(out[0], nominal[0])
    vin, vout = circuit.index_map["V(in)"], circuit.index_map["V(out)"]
    final = out[:, -1, :]
    assert float((final[:, vin] - 1.0).abs().max().item()) < 1e-12
    gap = final[:, vin] - final[:, vout]
    assert float(gap.min().item()) > 0.0 and float(gap.max().item()) < 0.02
    assert float((out[:, 1:, vout] - out[:, :-1, vout]).min().item()) >= -1e-13
