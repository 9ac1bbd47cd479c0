"""krets-result compatible writers: column names / order / values against the reference's checked-in goldens."""
import numpy as np
import pytest

import krets_b200 as kb
from krets_b200 import result_io


def _read(path):
    import pyarrow.parquet as pq
    t = pq.read_table(path)
    return t.column_names, {k: np.array(v) for k, v in t.to_pydict().items()}


def test_ac_writer_layout_matches_reference_golden(tmp_path, golden, oracle):
    """Feed the oracle's AC solution through the writer: the file must have the reference golden's columns, order and values."""
    g = golden["parquet"]["low_pass_filter_ac"]
    oc = oracle.Circuit(golden["netlists"]["low_pass_filter_ac"])
    s = g["spec"]
    fr, x = oc.solve_ac(oracle.generate_frequencies(s["sweep"], s["total_points"], s["fstart"], s["fstop"]))
    res = kb.AnalysisResult("ac", oc.names, x, fr)
    path = result_io.write_ac_results_to_parquet(res, str(tmp_path / "result"))
    assert path.endswith("result.parquet")
    names, cols = _read(path)
    assert names == ["frequency", "I(V1)_mag", "I(V1)_phase_deg", "V(in)_mag", "V(in)_phase_deg", "V(out)_mag", "V(out)_phase_deg"]
    for k in names:
        np.testing.assert_allclose(cols[k], g["columns"][k], rtol=1e-12, atol=1e-12, err_msg=k)


def test_dc_and_transient_writer_layout(tmp_path, golden, oracle):
    g = golden["parquet"]["inverter_dc"]
    oc = oracle.Circuit(golden["netlists"]["inverter"])
    s = g["spec"]
    rows = oc.solve_dc(s["element"], s["start"], s["stop"], s["step_size"])
    res = kb.AnalysisResult("dc", oc.names, rows, np.arange(rows.shape[0], dtype=np.float64))
    names, cols = _read(result_io.write_dc_results_to_parquet(res, str(tmp_path / "dc.parquet")))
    assert names == sorted(g["columns"].keys()) == ["I(V0)", "I(V1)", "V(D)", "V(G)", "V(S)", "step"]
    for k in names:
        np.testing.assert_allclose(cols[k], g["columns"][k], rtol=1e-12, atol=1e-15, err_msg=k)
    oc = oracle.Circuit(golden["netlists"]["dual_rc_ladder"])
    t, rows = oc.solve_tran(50e-6, 1e-3)
    res = kb.AnalysisResult("transient", oc.names, rows, t)
    names, cols = _read(result_io.write_tran_results_to_parquet(res, str(tmp_path / "tran")))
    assert names == ["time", "I(V1)", "V(in)", "V(int)", "V(out)"]
    assert np.array_equal(cols["time"], t) and np.array_equal(cols["V(out)"], rows[:, oc.index_map["V(out)"]])
    res = kb.AnalysisResult("op", oc.names, oc.solve_op()[None, :], None)
    names, cols = _read(result_io.write_op_results_to_parquet(res, str(tmp_path / "op")))
    assert names == sorted(oc.names) and all(len(cols[k]) == 1 for k in names)
