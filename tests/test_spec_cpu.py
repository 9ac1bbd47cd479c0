"""CPU tier: `krets.toml` analysis specs (crates/krets-parser/src/analyses.rs:6-78,161-165) and the circuit-path rule of
crates/krets-cli/src/main.rs:38-64, against the reference's own serde unit tests and its three checked-in krets.toml files."""
import os

import pytest

import krets_b200 as kb
from krets_b200 import spec


def _check(an, kind, want):
    assert an.kind == kind
    p = an.params
    if kind == "transient":
        assert p.time_step == want["time_step"] and p.stop_time == want["stop_time"]
    elif kind == "ac":
        assert (p.sweep.variation, p.sweep.points, p.fstart, p.fstop) == (want["variation"], want["points"], want["fstart"], want["fstop"])
    elif kind == "dc":
        assert (p.element, p.start, p.stop, p.step_size) == (want["element"], want["start"], want["stop"], want["step_size"])


def test_reference_serde_unit_tests(golden):
    """analyses.rs:170-268: the four TOML snippets of the reference's tests parse to the fields the reference asserts"""
    for text, what, kind, want, where in golden["assertions"]["toml"]:
        if what == "analysis":
            _check(spec.analysis_from_str(text), kind, want)
        else:
            s = spec.AnalysisSpec.from_str(text)
            assert s.circuit_path.endswith("krets.toml") and s.circuit_path == want["circuit_path"], where
            _check(s.analysis, kind, want)


def test_reference_krets_toml_files(golden):
    g = golden["specs"]
    s = spec.AnalysisSpec.from_str(g["inverter_dc"]["toml"])
    assert s.circuit_path == "inverter.cir"
    _check(s.analysis, "dc", golden["parquet"]["inverter_dc"]["spec"])
    s = spec.AnalysisSpec.from_str(g["low_pass_filter_ac"]["toml"])
    w = golden["parquet"]["low_pass_filter_ac"]["spec"]
    assert s.circuit_path == "ac.cir"
    _check(s.analysis, "ac", {"variation": w["sweep"], "points": w["total_points"], "fstart": w["fstart"], "fstop": w["fstop"]})
    s = spec.AnalysisSpec.from_str(g["high_pass_filter_transient"]["toml"])
    assert s.circuit_path == "transient.cir"
    _check(s.analysis, "transient", {"time_step": 10e-6, "stop_time": 2e-3})


def test_unit_variant_and_number_forms():
    assert spec.AnalysisSpec.from_str('circuit_path = "a.cir"\nanalysis = "op"\n').analysis.kind == "op"
    assert spec.AnalysisSpec.from_str('circuit_path = "a.cir"\n[analysis.op]\n').analysis.kind == "op"
    a = spec.analysis_from_str('[dc]\nelement = "V1"\nstart = 0\nstop = 5\nstep_size = 1\nignored = true\n')     # integers are numbers; unknown keys are ignored
    _check(a, "dc", {"element": "V1", "start": 0.0, "stop": 5.0, "step_size": 1.0})
    a = spec.analysis_from_str('[ac]\nsweep = { variation = "Octave", points_per_octave = 3 }\nfstart = 1\nfstop = 8\n')
    assert a.params.sweep.variation == "Octave" and a.params.sweep.points == 3
    assert kb.generate_frequencies(a.params.sweep.variation, 3, 1.0, 8.0)[-1] == 8.0


@pytest.mark.parametrize("text, fragment", [
    ('circuit_path = "a.cir"\n', "missing field `analysis`"),
    ('[analysis.transient]\ntime_step = 1e-6\nstop_time = 1e-3\n', "missing field `circuit_path`"),
    ('circuit_path = "a.cir"\n[analysis.transient]\ntime_step = 1e-6\n', "missing field `stop_time`"),
    ('circuit_path = "a.cir"\n[analysis.tran]\ntime_step = 1e-6\nstop_time = 1\n', "unknown variant `tran`"),
    ('circuit_path = "a.cir"\nanalysis = "noise"\n', "unknown variant `noise`"),
    ('circuit_path = "a.cir"\n[analysis.dc]\nelement = 1\nstart = 0\nstop = 1\nstep_size = 1\n', "expected a string"),
    ('circuit_path = "a.cir"\n[analysis.dc]\nelement = "V1"\nstart = "0"\nstop = 1\nstep_size = 1\n', "expected a number"),
    ('circuit_path = "a.cir"\n[analysis.ac]\nsweep = { variation = "Linear", total_points = -1 }\nfstart = 1\nfstop = 2\n', "expected u32"),
    ('circuit_path = "a.cir"\n[analysis.ac]\nsweep = { variation = "Linear", points_per_decade = 5 }\nfstart = 1\nfstop = 2\n', "missing field `total_points`"),
    ('circuit_path = "a.cir"\n[analysis.ac]\nsweep = { variation = "linear", total_points = 5 }\nfstart = 1\nfstop = 2\n', "unknown variant `linear`"),
    ('circuit_path = "a.cir"\n[analysis.ac]\nfstart = 1\nfstop = 2\n', "missing field `sweep`"),
    ('circuit_path = "a.cir"\n[analysis.dc]\nelement = "V1"\nstart = 0\nstop = 1\nstep_size = 1\n[analysis.op]\n', "exactly 1 element"),
    ('circuit_path = "a.cir"\n[analysis.Transient]\ntime_step = 1\nstop_time = 1\n', "unknown variant `Transient`"),      # rename_all = "lowercase"
    ('circuit_path = \n', "TOML parse error"),
])
def test_what_serde_rejects(text, fragment):
    with pytest.raises(spec.SpecError) as e:
        spec.AnalysisSpec.from_str(text)
    assert fragment in str(e.value)


def test_circuit_path_resolution(tmp_path):
    """main.rs:38-64: relative to the krets file first, else an existing absolute path, else an error naming both candidates"""
    d = tmp_path / "run"
    d.mkdir()
    (d / "a.cir").write_text("V1 in 0 1\nR1 in 0 1\n")
    other = tmp_path / "b.cir"
    other.write_text("V1 in 0 1\nR1 in 0 1\n")
    k = str(d / "krets.toml")
    assert spec.resolve_circuit_path(k, "a.cir") == str(d / "a.cir")
    assert spec.resolve_circuit_path(k, str(other)) == str(other)
    with pytest.raises(FileNotFoundError):
        spec.resolve_circuit_path(k, "b.cir")            # relative to the cwd is NOT tried (only krets-file-relative or absolute)


def _spec_dir(golden, key, tmp_path):
    s = golden["specs"][key]
    an = spec.AnalysisSpec.from_str(s["toml"])
    (tmp_path / an.circuit_path).write_text(golden["netlists"][s["netlist"]])
    k = tmp_path / "krets.toml"
    k.write_text(s["toml"])
    return str(k)


def test_run_spec_fails_loudly_without_a_device(golden, tmp_path):
    if kb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    k = _spec_dir(golden, "low_pass_filter_ac", tmp_path)
    with pytest.raises(kb.KretsError) as e:
        spec.run_spec(k)
    assert e.value.kind == "NoDevice" and not os.path.exists(tmp_path / "result.parquet")
    assert spec.main([k]) == 1 and spec.main([]) == 2
