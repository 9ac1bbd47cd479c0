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
