"""GPU tier: results computed by the CUDA path go through the krets-result compatible writers (krets_b200/result_io.py,
crates/krets-result/src/lib.rs:20-208) and the files match the reference's checked-in goldens column by column
(the writer tests used to feed oracle data only)."""
import numpy as np
import pytest

import krets_b200 as kb
from krets_b200 import result_io

pytestmark = pytest.mark.gpu


def _read(path):
    import pyarrow.parquet as pq
    t = pq.read_table(path)
    return t.column_names, {k: np.array(v) for k, v in t.to_pydict().items()}


def test_cuda_ac_result_file_equals_reference_golden(tmp_path, golden):
    """circuits/low_pass_filter/ac: krets.toml spec -> Solver (CUDA) -> write_ac_results_to_parquet == circuits/low_pass_filter/ac/result.parquet"""
    assert kb.device_count() > 0
    g = golden["parquet"]["low_pass_filter_ac"]
    s = g["spec"]
    circuit = kb.parse_circuit_description(golden["netlists"]["low_pass_filter_ac"])
    before = kb.kernel_launch_count()
    res = kb.Solver(circuit).solve(kb.Analysis.Ac(kb.AcAnalysis(kb.AcSweep(s["sweep"], s["total_points"]), s["fstart"], s["fstop"])))
    assert kb.kernel_launch_count() > before
    names, cols = _read(result_io.write_ac_results_to_parquet(res, str(tmp_path / "result")))
    assert names == ["frequency", "I(V1)_mag", "I(V1)_phase_deg", "V(in)_mag", "V(in)_phase_deg", "V(out)_mag", "V(out)_phase_deg"]
    assert np.array_equal(cols["frequency"], np.array(g["columns"]["frequency"]))          # the grid is bit-exact
    for k in names[1:]:
        np.testing.assert_allclose(cols[k], g["columns"][k], rtol=1e-9, atol=1e-12, err_msg=k)


def test_cuda_dc_and_transient_result_files(tmp_path, golden, oracle):
    """inverter DC sweep (NMOS stamps nothing, SURVEY B-7) against circuits/inverter/dc/result.parquet; dual_rc_ladder transient
    against the oracle's rows, through the writers' column rules (`step` sorted in, `time` first)."""
    g = golden["parquet"]["inverter_dc"]
    s = g["spec"]
    circuit = kb.parse_circuit_description(golden["netlists"]["inverter"])
    res = kb.Solver(circuit).solve(kb.Analysis.Dc(kb.DcAnalysis(s["element"], s["start"], s["stop"], s["step_size"])))
    names, cols = _read(result_io.write_dc_results_to_parquet(res, str(tmp_path / "dc.parquet")))
    assert names == sorted(g["columns"].keys())
    for k in names:
        np.testing.assert_allclose(cols[k], g["columns"][k], rtol=1e-9, atol=1e-15, err_msg=k)
    text = golden["netlists"]["dual_rc_ladder"]
    circuit, oc = kb.parse_circuit_description(text), oracle.Circuit(text)
    res = kb.Solver(circuit).solve(kb.Analysis.Transient(kb.TransientAnalysis(50e-6, 1e-3)))
    t, rows = oc.solve_tran(50e-6, 1e-3)
    names, cols = _read(result_io.write_tran_results_to_parquet(res, str(tmp_path / "tran")))
    assert names == ["time", "I(V1)", "V(in)", "V(int)", "V(out)"]
    assert np.array_equal(cols["time"], t)
    for k in names[1:]:
        np.testing.assert_allclose(cols[k], rows[:, oc.index_map[k]], rtol=1e-9, atol=1e-12, err_msg=k)


# ---- `krets <krets.toml>` (crates/krets-cli/src/main.rs:33-125) on the reference's three checked-in specs --------------------------
def _spec_dir(golden, key, tmp_path):
    from krets_b200 import spec
    s = golden["specs"][key]
    an = spec.AnalysisSpec.from_str(s["toml"])
    d = tmp_path / key
    d.mkdir()
    (d / an.circuit_path).write_text(golden["netlists"][s["netlist"]])
    (d / "krets.toml").write_text(s["toml"])
    return str(d / "krets.toml"), d


def test_run_spec_writes_the_reference_result_files(tmp_path, golden, oracle):
    """circuits/low_pass_filter/ac/krets.toml and circuits/inverter/dc/krets.toml run like `krets <file>`: result.parquet next to the spec
    equals the result.parquet the reference checked in; circuits/high_pass_filter/transient/krets.toml against the oracle's rows."""
    from krets_b200 import spec
    k, d = _spec_dir(golden, "low_pass_filter_ac", tmp_path)
    res, out = spec.run_spec(k)
    assert res.kind == "ac" and out == str(d / "result.parquet")
    g = golden["parquet"]["low_pass_filter_ac"]
    names, cols = _read(out)
    assert names == ["frequency", "I(V1)_mag", "I(V1)_phase_deg", "V(in)_mag", "V(in)_phase_deg", "V(out)_mag", "V(out)_phase_deg"]
    assert np.array_equal(cols["frequency"], np.array(g["columns"]["frequency"]))
    for c in names[1:]:
        np.testing.assert_allclose(cols[c], g["columns"][c], rtol=1e-9, atol=1e-12, err_msg=c)

    k, d = _spec_dir(golden, "inverter_dc", tmp_path)
    assert spec.main([k]) == 0                       # the module's entry point: same run, exit status 0
    g = golden["parquet"]["inverter_dc"]
    names, cols = _read(str(d / "result.parquet"))
    assert names == sorted(g["columns"].keys())
    for c in names:
        np.testing.assert_allclose(cols[c], g["columns"][c], rtol=1e-9, atol=1e-15, err_msg=c)

    k, d = _spec_dir(golden, "high_pass_filter_transient", tmp_path)
    res, out = spec.run_spec(k)
    oc = oracle.Circuit(golden["netlists"][golden["specs"]["high_pass_filter_transient"]["netlist"]])
    t, rows = oc.solve_tran(10e-6, 2e-3)
    names, cols = _read(out)
    assert names[0] == "time" and names[1:] == sorted(names[1:]) and np.array_equal(cols["time"], t) and len(t) == 201
    for c in names[1:]:
        np.testing.assert_allclose(cols[c], rows[:, oc.index_map[c]], rtol=1e-9, atol=1e-12, err_msg=c)
