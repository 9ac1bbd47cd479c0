"""krets-result compatible Parquet writers (crates/krets-result/src/lib.rs:20-208).

Same file layout as the reference so that krets-gui and diffs against checked-in goldens keep working:
  OP         one row, one f64 column per unknown (HashMap order in the reference: here sorted, order is not part of the contract)
  DC         columns sorted lexicographically (includes the `step` column)                              lib.rs:54-88
  transient  `time` first, then the remaining columns sorted                                           lib.rs:90-129
  AC         `frequency`, then for every key K (sorted) `K_mag` = sqrt(re^2+im^2) and `K_phase_deg` = atan2(im, re)*180/PI   lib.rs:140-208
Host-side I/O only (pyarrow); nothing here touches the solver path.
"""
from __future__ import annotations

import math
import os

import numpy as np


def _ensure_parquet_extension(filename: str) -> str:      # lib.rs:10-17
    return filename if os.path.splitext(filename)[1] == ".parquet" else f"{filename}.parquet"


def _write(columns, filename):
    import pyarrow as pa
    import pyarrow.parquet as pq
    table = pa.table({name: pa.array(np.asarray(vals, dtype=np.float64)) for name, vals in columns})
    path = _ensure_parquet_extension(filename)
    pq.write_table(table, path)
    return path


def write_op_results_to_parquet(result, filename: str):
    """result: AnalysisResult of kind 'op'."""
    row = result.into_op()
    return _write([(k, [row[k]]) for k in sorted(row)], filename)


def write_dc_results_to_parquet(result, filename: str):
    if result.array.shape[0] == 0:
        return None
    cols = {nm: result.array[:, i] for i, nm in enumerate(result.names)}
    cols["step"] = result.axis
    return _write([(k, cols[k]) for k in sorted(cols)], filename)


def write_tran_results_to_parquet(result, filename: str):
    if result.array.shape[0] == 0:
        return None
    cols = {nm: result.array[:, i] for i, nm in enumerate(result.names)}
    order = ["time"] + sorted(cols)
    cols["time"] = result.axis
    return _write([(k, cols[k]) for k in order], filename)


def write_ac_results_to_parquet(result, filename: str):
    if result.array.shape[0] == 0:
        return None
    out = [("frequency", result.axis)]
    for i in sorted(range(len(result.names)), key=lambda j: result.names[j]):
        v = result.array[:, i]
        mag = np.sqrt(v.real * v.real + v.imag * v.imag)
        phase = np.arctan2(v.imag, v.real) * 180.0 / math.pi
        out.append((f"{result.names[i]}_mag", mag))
        out.append((f"{result.names[i]}_phase_deg", phase))
    return _write(out, filename)
