"""`krets.toml` analysis specs and the call order of `krets <krets.toml>` — the caller's side of the hot path.

Restates, over the CUDA backend:
  * `AnalysisSpec` / `AnalysisSpec::from_file`            crates/krets-parser/src/analyses.rs:6-24
  * the serde shape of `Analysis`, `DcAnalysis`, `AcSweep`, `AcAnalysis`, `TransientAnalysis`   analyses.rs:26-78,161-165
    (externally tagged enum with lowercase variant names: `[analysis.dc]`, `[analysis.ac]`, `[analysis.transient]`, `analysis = "op"`;
    `AcSweep` internally tagged by `variation`; unknown keys are ignored, missing ones are errors — serde defaults)
  * the run of crates/krets-cli/src/main.rs:33-125: circuit path resolved relative to the krets file first, then as an absolute
    path; `SolverConfig::default()`; `Solver::solve`; `result.parquet` written next to the krets file by the writer of the result kind.
The GUI launch (main.rs:127-136), logging and argument parsing are out of scope.

    python -m krets_b200.spec circuits/low_pass_filter/ac/krets.toml
"""
from __future__ import annotations

import os
import sys
import tomllib
from dataclasses import dataclass

from . import result_io
from .api import AcAnalysis, AcSweep, Analysis, AnalysisResult, DcAnalysis, Solver, SolverConfig, TransientAnalysis, parse_circuit_description_file


class SpecError(ValueError):
    """what `toml::from_str::<AnalysisSpec>` rejects (analyses.rs:21): TOML syntax, missing fields, wrong types, unknown variants"""


def _f64(table, key, where):
    if key not in table:
        raise SpecError(f"missing field `{key}` in {where}")
    v = table[key]
    if isinstance(v, bool) or not isinstance(v, (int, float)):          # serde's f64 visitor takes TOML integers and floats
        raise SpecError(f"invalid type for `{key}` in {where}: expected a number, found {type(v).__name__}")
    return float(v)


def _u32(table, key, where):
    if key not in table:
        raise SpecError(f"missing field `{key}` in {where}")
    v = table[key]
    if isinstance(v, bool) or not isinstance(v, int) or not 0 <= v <= 0xFFFFFFFF:
        raise SpecError(f"invalid value for `{key}` in {where}: expected u32, found {v!r}")
    return v


def _string(table, key, where):
    if key not in table:
        raise SpecError(f"missing field `{key}` in {where}")
    if not isinstance(table[key], str):
        raise SpecError(f"invalid type for `{key}` in {where}: expected a string")
    return table[key]


def _table(v, where):
    if not isinstance(v, dict):
        raise SpecError(f"invalid type in {where}: expected a table")
    return v


_SWEEP_FIELD = {"Decade": "points_per_decade", "Octave": "points_per_octave", "Linear": "total_points"}     # analyses.rs:57-65


def ac_sweep_from_toml(v) -> AcSweep:
    t = _table(v, "sweep")
    variation = _string(t, "variation", "sweep")                        # #[serde(tag = "variation")]
    if variation not in _SWEEP_FIELD:
        raise SpecError(f"unknown variant `{variation}`, expected one of `Decade`, `Octave`, `Linear`")
    return AcSweep(variation, _u32(t, _SWEEP_FIELD[variation], f"sweep ({variation})"))


def analysis_from_toml(v) -> Analysis:
    """`Analysis` as serde reads it (analyses.rs:27-41): the string "op", or a table with exactly one lowercase variant key."""
    if isinstance(v, str):
        if v == "op":
            return Analysis.Op()
        if v in ("dc", "ac", "transient"):
            raise SpecError(f"invalid type: unit variant, expected struct variant `{v}`")
        raise SpecError(f"unknown variant `{v}`, expected one of `op`, `dc`, `ac`, `transient`")
    t = _table(v, "analysis")
    if len(t) != 1:
        raise SpecError(f"wanted exactly 1 element, found {len(t)} elements")
    (kind, body), = t.items()
    if kind == "op":
        if isinstance(body, dict) and not body:
            return Analysis.Op()
        raise SpecError("expected an empty table for the unit variant `op`")
    if kind == "dc":                                                    # analyses.rs:44-54
        b = _table(body, "dc")
        return Analysis.Dc(DcAnalysis(_string(b, "element", "dc"), _f64(b, "start", "dc"), _f64(b, "stop", "dc"), _f64(b, "step_size", "dc")))
    if kind == "ac":                                                    # analyses.rs:68-77
        b = _table(body, "ac")
        if "sweep" not in b:
            raise SpecError("missing field `sweep` in ac")
        return Analysis.Ac(AcAnalysis(ac_sweep_from_toml(b["sweep"]), _f64(b, "fstart", "ac"), _f64(b, "fstop", "ac")))
    if kind == "transient":                                             # analyses.rs:161-165
        b = _table(body, "transient")
        return Analysis.Transient(TransientAnalysis(_f64(b, "time_step", "transient"), _f64(b, "stop_time", "transient")))
    raise SpecError(f"unknown variant `{kind}`, expected one of `op`, `dc`, `ac`, `transient`")


def analysis_from_str(text: str) -> Analysis:
    """`toml::from_str::<Analysis>` (what analyses.rs:173-187, :218-240 exercise)"""
    try:
        return analysis_from_toml(tomllib.loads(text))
    except tomllib.TOMLDecodeError as e:
        raise SpecError(f"TOML parse error: {e}") from None


@dataclass
class AnalysisSpec:                                                     # analyses.rs:6-13
    circuit_path: str
    analysis: Analysis

    @staticmethod
    def from_str(text: str) -> "AnalysisSpec":
        try:
            t = tomllib.loads(text)
        except tomllib.TOMLDecodeError as e:
            raise SpecError(f"TOML parse error: {e}") from None
        path = _string(t, "circuit_path", "the spec")
        if "analysis" not in t:
            raise SpecError("missing field `analysis` in the spec")
        return AnalysisSpec(path, analysis_from_toml(t["analysis"]))

    @staticmethod
    def from_file(path: str) -> "AnalysisSpec":                         # analyses.rs:19-23
        with open(path, "r", encoding="utf-8") as f:
            return AnalysisSpec.from_str(f.read())


def resolve_circuit_path(krets_file: str, circuit_path: str) -> str:
    """main.rs:38-64: relative to the krets file first; else the path as given if it is absolute and exists."""
    parent = os.path.dirname(krets_file) or "."
    candidate = os.path.join(parent, circuit_path)
    if os.path.exists(candidate):
        return candidate
    if os.path.isabs(circuit_path) and os.path.exists(circuit_path):
        return circuit_path
    raise FileNotFoundError(f"Circuit file not found. Tried: relative to krets file: {candidate}; as given: {circuit_path}")


_WRITERS = {"op": result_io.write_op_results_to_parquet, "dc": result_io.write_dc_results_to_parquet,
            "ac": result_io.write_ac_results_to_parquet, "transient": result_io.write_tran_results_to_parquet}     # main.rs:98-123


def run_spec(krets_file: str, *, config: SolverConfig | None = None, write: bool = True) -> tuple[AnalysisResult, str | None]:
    """What `krets <krets_file>` does up to the GUI (main.rs:33-125), the solve on the GPU: returns the result and the path of the
    `result.parquet` written next to the krets file (None with write=False)."""
    spec = AnalysisSpec.from_file(krets_file)
    circuit = parse_circuit_description_file(resolve_circuit_path(krets_file, spec.circuit_path))      # main.rs:67
    result = Solver(circuit, config or SolverConfig.default()).solve(spec.analysis)                      # main.rs:80-93
    out = None
    if write:
        out = _WRITERS[result.kind](result, os.path.join(os.path.dirname(krets_file) or ".", "result.parquet"))   # main.rs:45-46
    return result, out


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage: python -m krets_b200.spec <krets.toml>", file=sys.stderr)
        return 2
    try:
        result, out = run_spec(argv[0])
    except Exception as e:            # main.rs exits with status 1 on every failure of the run
        print(f"Error: {e}", file=sys.stderr)
        return 1
    print(f"{result.kind}: {result.array.shape[0]} row(s) x {result.array.shape[1]} unknown(s) -> {out}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
