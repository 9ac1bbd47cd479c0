#!/usr/bin/env python
"""BASELINE config 5 (RLC mesh, Monte-Carlo lanes) on one GPU through the front executor: device-resident inputs and
outputs, CUDA events.  python tools/measure_mesh.py [--grid 224] [--lanes 128] [--steps 16] [--modes tran,ac,dc]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import netgen
import krets_b200 as kb
from krets_b200 import api


def mc_matrix(circuit, varied, lanes, seed=0x6B72657473):
    """values[k][lane] = nominal * (0.95 + 0.10 u), u from splitmix64(seed + n_elements*lane + element); lane 0 nominal."""
    e = np.asarray(varied, dtype=np.uint64)[:, None]
    lane = np.arange(lanes, dtype=np.uint64)[None, :]
    x = (np.uint64(seed) + np.uint64(circuit.num_elements) * lane + e)
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    f = 0.95 + 0.10 * u
    f[:, 0] = 1.0
    nominal = np.array([circuit.element_value(int(k)) for k in varied])[:, None]
    return nominal * f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=224)
    ap.add_argument("--lanes", type=int, default=128)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--modes", default="tran,ac")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    peak = 6453.1
    res = []
    for mode in args.modes.split(","):
        transient = mode != "ac"
        text = netgen.rlc_mesh(args.grid, args.grid, transient=transient)
        circuit = kb.parse_circuit_description(text)
        kinds = [circuit.element(e).kind for e in range(circuit.num_elements)]
        varied = [e for e, k in enumerate(kinds) if k in (api.KIND_R, api.KIND_C, api.KIND_L)]
        vals = torch.from_numpy(mc_matrix(circuit, varied, args.lanes)).cuda()
        t0 = time.time()
        if mode == "ac":
            plan = kb.Plan(circuit, api.MODE_AC, varied=varied, nominal_frequency=1e6, executor=api.EXEC_STAGED)
        else:
            plan_dc = kb.Plan(circuit, api.MODE_DC, varied=varied, executor=api.EXEC_STAGED)
            plan = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=1e-9, integrator=api.INTEGRATOR_TRAP, executor=api.EXEC_STAGED) if mode == "tran" else plan_dc
        t_an = time.time() - t0
        ft = plan.front_tables()
        lanes = args.lanes
        if mode == "ac":
            freqs = torch.full((lanes,), 1e6, dtype=torch.float64, device="cuda") * torch.logspace(-2, 2, lanes, dtype=torch.float64, device="cuda")
            out = torch.empty((lanes, circuit.n), dtype=torch.complex128, device="cuda")
            fn = lambda: plan.solve_ac(freqs, varied_values=vals, out=out)
            points = lanes
        elif mode == "dc":
            out = torch.empty((lanes, circuit.n), dtype=torch.float64, device="cuda")
            fn = lambda: plan.solve_op(lanes, varied_values=vals, out=out)
            points = lanes
        else:
            out = torch.empty((lanes, args.steps + 1, circuit.n), dtype=torch.float64, device="cuda")
            fn = lambda: plan.solve_transient(plan_dc, args.steps, lanes, varied_values=vals, out=out)
            points = lanes * (args.steps + 1)
        _, status = fn()
        torch.cuda.synchronize()
        bad = int((status != 0).sum().item())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = kb.kernel_launch_count()
        e0.record()
        for _ in range(args.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        launches = (kb.kernel_launch_count() - l0) // args.reps
        s = 16 if mode == "ac" else 8
        rec = {"config": 5, "mode": mode, "grid": args.grid, "n": circuit.n, "nnz": int(plan.info.nnz), "nnz_lu": int(plan.info.nnz_lu),
               "fronts": ft["n_fronts"], "levels": ft["n_levels"], "max_front": ft["max_m"], "factor_flops": ft["flops"] * (4 if mode == "ac" else 1),
               "lanes": lanes, "steps": args.steps if mode == "tran" else 0, "analyze_s": round(t_an, 2), "ms": ms, "points_per_s": points / (ms * 1e-3),
               "launches": int(launches), "lanes_bad_status": bad, "finite": bool(torch.isfinite(torch.view_as_real(out) if mode == "ac" else out).all().item()),
               "alg_bytes_per_point": int(plan.info.alg_bytes_per_point)}
        if mode == "tran":
            # first row: DC factor + solve, second: transient factor; every further step re-reads the factor once (SURVEY §8d B_step)
            step_bytes = s * (ft["lu_size"] + 3 * circuit.n)
            rec["step_bytes_per_point"] = int(step_bytes)
        else:
            rec["roofline_frac_alg"] = rec["alg_bytes_per_point"] * rec["points_per_s"] / 1e9 / peak
        res.append(rec)
        print(json.dumps(rec), flush=True)
        del plan
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "mesh_r1.jsonl"), "a") as f:
        for r in res:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
