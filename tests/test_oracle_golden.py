"""Pin the CPU oracle against everything the reference's own tests hold for the hot path.

Oracle = test infrastructure.  These are the `-m "not gpu"` checks that make the oracle trustworthy before it is
used to judge the CUDA path.
"""
import math

import numpy as np
import pytest

import netgen


def circ(oracle, golden, key):
    if key.startswith("resistor_ladder_"):
        return oracle.Circuit(netgen.resistor_ladder(int(key.rsplit("_", 1)[1])))
    return oracle.Circuit(golden["netlists"][key])


# ---- index map (parser.rs:105-125) bit-exact -------------------------------------------------------------
def test_index_map_known(oracle, golden):
    assert circ(oracle, golden, "low_pass_filter").index_map == {"I(V1)": 0, "V(in)": 1, "V(out)": 2}
    assert circ(oracle, golden, "dual_rc_ladder").index_map == {"V(int)": 0, "V(in)": 1, "I(V1)": 2, "V(out)": 3}
    c = circ(oracle, golden, "resistor_ladder_500")
    assert c.n == 501 and c.index_map["I(V1)"] == 0 and all(c.index_map[f"V({k})"] == k for k in range(1, 501))
    # inverter: MN1 D G S 0 NMOS first -> V(D),V(G),V(S); then V0 -> I(V0); V1 -> I(V1)
    assert circ(oracle, golden, "inverter").index_map == {"V(D)": 0, "V(G)": 1, "V(S)": 2, "I(V0)": 3, "I(V1)": 4}
    # book circuit: 13 elements -> 8 nodes + 3 V + 2 I = 13 unknowns (test_parse_netlist.rs: 9 nodes incl. ground)
    c = circ(oracle, golden, "farid_n_najm")
    assert c.num_elements == 13 and c.n == 13


def test_empty_netlist(oracle):
    with pytest.raises(oracle.OracleError) as e:
        oracle.Circuit("* nothing\n")
    assert e.value.kind == "EmptyNetlist"


def test_value_suffixes(oracle):
    # utils.rs:94-110
    c = oracle.Circuit("R1 a 0 1.5k\nC1 a 0 10u\nC2 a 0 22n\nC3 a 0 1.2p\nR2 a 0 3MEG\nR3 a 0 100\nC4 a 0 1e-6\nC5 a 0 1.23E-9\n")
    vals = [c.element_value(i) for i in range(8)]
    exp = [1500.0, 10e-6, 22e-9, 1.2e-12, 3e6, 100.0, 1e-6, 1.23e-9]
    assert all(abs(a - b) < 1e-15 * max(1, abs(b)) for a, b in zip(vals, exp))
    for bad in ("R1 a 0 1.5x\n", "R1 a 0 garbage\n"):
        with pytest.raises(oracle.OracleError):
            oracle.Circuit(bad)


def test_waveforms(oracle):
    # voltage_source.rs:370-423
    c = oracle.Circuit("V1 in 0 PULSE(0 5 1u 10n 10n 5u 10u)\nR1 in 0 1\n")
    for t, v in [(0.5e-6, 0.0), (1.005e-6, 2.5), (3.01e-6, 5.0), (6.015e-6, 2.5), (8.0e-6, 0.0), (11.005e-6, 2.5)]:
        assert abs(c.transient_value_at(0, t) - v) < 1e-9
    # voltage_source.rs:445-496
    c = oracle.Circuit("V1 in 0 SIN(1 5 1k 1m 0 0)\nR1 in 0 1\n")
    for t, v in [(0.5e-3, 1.0), (1.25e-3, 6.0), (1.5e-3, 1.0), (1.75e-3, -4.0)]:
        assert abs(c.transient_value_at(0, t) - v) < 1e-9
    c = oracle.Circuit("V1 in 0 SIN(1 5 1k 1m 0 90)\nR1 in 0 1\n")
    assert abs(c.transient_value_at(0, 1e-3) - 6.0) < 1e-9


# ---- OP (test_op.rs) ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("key", ["farid_n_najm", "test_case_1", "voltage_divider", "low_pass_filter", "high_pass_filter",
                                 "basic_001", "resistor_ladder_500", "diode_iv_curve"])
def test_op_assertions(oracle, golden, key):
    c = circ(oracle, golden, key)
    x = c.solve_op()
    for name, exp, tol, where in golden["assertions"]["op"][key]:
        assert abs(x[c.index_map[name]] - exp) < tol, (name, where)


def test_diode_op_iterations(oracle, golden):
    # SURVEY.md §B-6 [probe]: converges in 4 iterations to 0.51693 V
    c = circ(oracle, golden, "diode_iv_curve")
    x, it = c.solve_op(return_iterations=True)
    assert abs(x[c.index_map["V(out)"]] - 0.51693) < 2e-5


# ---- DC (test_dc.rs) ---------------------------------------------------------------------------------------
def test_dc_assertions(oracle, golden):
    a = golden["assertions"]["dc"]["voltage_divider"]
    c = circ(oracle, golden, "voltage_divider")
    rows = c.solve_dc(a["element"], a["start"], a["stop"], a["step_size"])
    assert rows.shape[0] == a["len"]
    for r, checks in a["rows"].items():
        for name, exp, tol, where in checks:
            assert abs(rows[int(r)][c.index_map[name]] - exp) < tol, (name, where)
    # the swept value is restored (dc.rs:117-121)
    assert c.element_value(c.element_index("V1")) == 1.0


def test_dc_errors(oracle, golden):
    c = circ(oracle, golden, "voltage_divider")
    with pytest.raises(oracle.OracleError) as e:
        c.solve_dc("V9", 0, 1, 1)
    assert e.value.kind == "ElementNotFound"
    with pytest.raises(oracle.OracleError) as e:
        c.solve_dc("R1", 0, 1, 1)
    assert e.value.kind == "InvalidElementFormat"


def test_dc_golden_inverter(oracle, golden):
    """circuits/inverter/dc/result.parquet: 101 rows; pins the NMOS no-op and the step column."""
    g = golden["parquet"]["inverter_dc"]
    c = circ(oracle, golden, "inverter")
    s = g["spec"]
    rows = c.solve_dc(s["element"], s["start"], s["stop"], s["step_size"])
    cols = g["columns"]
    assert rows.shape[0] == len(cols["step"]) == 101
    for name, col in cols.items():
        if name == "step":
            assert col == [float(i) for i in range(101)]
            continue
        got = rows[:, c.index_map[name]]
        np.testing.assert_allclose(got, np.array(col), rtol=1e-12, atol=1e-15, err_msg=name)


# ---- AC (test_ac.rs + golden parquet) ---------------------------------------------------------------------
@pytest.mark.parametrize("key", ["low_pass_filter", "high_pass_filter"])
def test_ac_assertions(oracle, golden, key):
    c = circ(oracle, golden, key)
    f = oracle.generate_frequencies("Linear", 1, 1000.0, 1000.0)
    fr, x = c.solve_ac(f)
    for name, re, im, tol, where in golden["assertions"]["ac"][key]:
        v = complex(fr[0], 0.0) if name == "frequency" else x[0][c.index_map[name]]
        assert abs(v.real - re) < tol and abs(v.imag - im) < tol, (name, where)


def test_ac_closed_form(oracle, golden):
    # test_ac.rs:87-104
    c = circ(oracle, golden, "low_pass_filter")
    fr, x = c.solve_ac(oracle.generate_frequencies("Linear", 100, 1.0, 1000.0))
    for f, row in zip(fr, x):
        w = 2 * math.pi * f * 1000.0 * 1e-6
        v = row[c.index_map["V(out)"]]
        assert abs(v.real - 1 / (1 + w * w)) < 1e-3 and abs(v.imag + w / (1 + w * w)) < 1e-3


def test_ac_golden_parquet(oracle, golden):
    """Full-precision krets+faer output: circuits/low_pass_filter/ac/result.parquet (10 points)."""
    g = golden["parquet"]["low_pass_filter_ac"]
    c = circ(oracle, golden, "low_pass_filter_ac")
    s = g["spec"]
    freqs = oracle.generate_frequencies(s["sweep"], s["total_points"], s["fstart"], s["fstop"])
    fr, x = c.solve_ac(freqs)
    cols = g["columns"]
    assert list(fr) == cols["frequency"]                      # bit-exact frequency grid (analyses.rs:97-99)
    for name in ("I(V1)", "V(in)", "V(out)"):
        v = x[:, c.index_map[name]]
        mag = np.abs(v)
        ph = np.degrees(np.arctan2(v.imag, v.real))          # krets-result lib.rs:166-200
        np.testing.assert_allclose(mag, cols[name + "_mag"], rtol=2e-15, atol=0, err_msg=name)
        np.testing.assert_allclose(ph, cols[name + "_phase_deg"], rtol=2e-13, atol=1e-13, err_msg=name)


def test_ac_current_source_singular(oracle, golden):
    # AC with an I source has an empty row -> DecompositionFailed (ac.rs:70)
    c = circ(oracle, golden, "basic_001")
    with pytest.raises(oracle.OracleError) as e:
        c.solve_ac([1000.0])
    assert e.value.kind == "DecompositionFailed"


# ---- frequency grids (analyses.rs:80-158) ---------------------------------------------------------------
def test_frequency_grids(oracle):
    f = oracle.generate_frequencies("Linear", 10, 0.1, 1000.0)
    assert f[0] == 0.1 and f[1] == 0.1 + 1 * ((1000.0 - 0.1) / 9)
    d = oracle.generate_frequencies("Decade", 10, 1.0, 1e6)
    assert len(d) == 61 and d[0] == 1.0 and abs(d[-1] / 1e6 - 1) < 1e-9 and np.all(np.diff(d) > 0)
    # the recurrence is sequential multiplication, not pow(i): drift is part of the input
    fac = 10.0 ** (1.0 / 10)
    cur = 1.0
    for i in range(5):
        assert d[i] == cur
        cur *= fac
    o = oracle.generate_frequencies("Octave", 3, 10.0, 80.0)
    assert len(o) == 10 and o[0] == 10.0
    assert len(oracle.generate_frequencies("Decade", 10, 0.0, 10.0)) == 0
    assert len(oracle.generate_frequencies("Linear", 0, 1.0, 10.0)) == 0
    assert list(oracle.generate_frequencies("Linear", 1, 5.0, 10.0)) == [5.0]


# ---- transient (test_transient.rs) --------------------------------------------------------------------
@pytest.mark.parametrize("key", ["dual_rc_ladder", "low_pass_filter_transient", "high_pass_filter_transient"])
def test_transient_assertions(oracle, golden, key):
    a = golden["assertions"]["transient"][key]
    c = circ(oracle, golden, key)
    t, rows = c.solve_tran(a["time_step"], a["stop_time"])
    for r, checks in a["rows"].items():
        for name, exp, tol, where in checks:
            got = t[int(r)] if name == "time" else rows[int(r)][c.index_map[name]]
            assert abs(got - exp) < tol, (name, where, got)


def test_transient_source_lag(oracle, golden):
    """SURVEY.md §B-1 [probe]: row 42 of the low-pass transient is 0.86472 with the one-step source lag."""
    c = circ(oracle, golden, "low_pass_filter_transient")
    t, rows = c.solve_tran(50e-6, 20e-3)
    assert len(t) == 401
    assert abs(rows[42][c.index_map["V(out)"]] - 0.86472) < 2e-5


def test_rectifier_runs(oracle, golden):
    # test_transient.rs:164-178: nonlinear transient runs to completion
    c = circ(oracle, golden, "rectifier")
    t, rows = c.solve_tran(50e-6, 50e-3)
    assert len(t) == 1001 and np.all(np.isfinite(rows))
    assert rows[:, c.index_map["V(out_dc)"]].max() > 3.5      # rectified ~5 V peak minus a diode drop


# ---- second opinion on the LU (scipy SuperLU), SURVEY.md §8c --------------------------------------------
def test_lu_vs_scipy(oracle):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    c = oracle.Circuit(netgen.rc_ladder(200))
    colptr, rowidx, vals = c.stamp_matrix("ac", 12345.0)
    A = sp.csc_matrix((vals, rowidx, colptr), shape=(c.n, c.n))
    b = np.zeros(c.n, dtype=complex)
    b[c.index_map["I(V1)"]] = 1.0
    ref = spla.splu(A).solve(b)
    fr, x = c.solve_ac([12345.0])
    np.testing.assert_allclose(x[0], ref, rtol=1e-10, atol=1e-16)
    c = oracle.Circuit(netgen.rlc_mesh(6, 7, transient=True))
    colptr, rowidx, vals = c.stamp_matrix("dc")
    A = sp.csc_matrix((vals, rowidx, colptr), shape=(c.n, c.n))
    b = np.zeros(c.n)
    b[c.index_map["I(I1)"]] = 1e-3
    ref = spla.splu(A).solve(b)
    np.testing.assert_allclose(c.solve_op(), ref, rtol=1e-9, atol=1e-14)


def test_cpu_best_effort_mode_is_bit_identical(oracle):
    """the tuned CPU mode (analysis once, re-stamp + numeric LU per point) sums every entry in the same order as
    sum_triplets, so it must reproduce the faithful restatement bit for bit."""
    import netgen
    oc = oracle.Circuit(netgen.resistor_ladder(300))
    a = oc.solve_dc("V1", 0.0, 63 * 2.0 ** -6, 2.0 ** -6)
    b = oc.solve_dc_best_effort("V1", 0.0, 63 * 2.0 ** -6, 2.0 ** -6, nthreads=3)
    assert a.shape == b.shape == (64, 301) and np.array_equal(a, b)
