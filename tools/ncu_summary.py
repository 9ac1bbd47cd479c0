#!/usr/bin/env python
"""Summarise an `ncu --set full` report (.ncu-rep) into the JSON kept under profiles/: the metrics DESIGN.md argues from.
python tools/ncu_summary.py report.ncu-rep out.json "workload" "command" [lanes]"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(rep, out, workload, command, lanes=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, recs = rows[0], rows[1], rows[2:]
    res = []
    for r in recs:
        m = {k: {"unit": units[head.index(k)], "value": r[head.index(k)]} for k in KEYS if k in head}
        rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        total = float(rd["value"].replace(",", "")) * SCALE[rd["unit"]] + float(wr["value"].replace(",", "")) * SCALE[wr["unit"]]
        d = {"kernel": r[head.index("Kernel Name")], "workload": workload, "command": command, "metrics": m, "dram_bytes_per_launch": total}
        if lanes:
            d["dram_bytes_per_point"] = total / float(lanes)
        res.append(d)
    with open(out, "w") as f:
        json.dump(res[0] if len(res) == 1 else res, f, indent=1)
    print(json.dumps({"kernels": [d["kernel"] for d in res], "dram_bytes_per_launch": [d["dram_bytes_per_launch"] for d in res]}))


if __name__ == "__main__":
    main(*sys.argv[1:6])
