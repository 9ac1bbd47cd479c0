#!/usr/bin/env python
"""Measure the other BASELINE.json configs on one GPU (device-resident, CUDA events) — numbers for profiles/, not the
driver's bench line.  python tools/measure_configs.py [--quick]"""
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


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    quick = "--quick" in sys.argv
    out = []
    peak = 6453.1
    # config 1: low_pass_filter AC, 100 linear points 1 Hz..1 kHz
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        golden = json.load(f)
    c = kb.parse_circuit_description(golden["netlists"]["low_pass_filter"])
    plan = kb.Plan(c, api.MODE_AC)
    fr = torch.from_numpy(kb.generate_frequencies("Linear", 100, 1.0, 1000.0)).cuda()
    ms = timed(lambda: plan.solve_ac(fr), reps=20, warm=3)
    out.append({"config": 1, "what": "low_pass_filter AC 100 pts (n=3, c64)", "ms": ms, "points_per_s": 100 / (ms * 1e-3),
                "executor": plan.info.executor, "note": "latency/launch bound: 100 lanes = 4 warps"})
    # config 3: RC ladder 5000 AC, decade grid 1 Hz..1e8 Hz, 125000 points/decade
    c = kb.parse_circuit_description(netgen.rc_ladder(5000))
    plan = kb.Plan(c, api.MODE_AC, nominal_frequency=1e4)
    allf = kb.generate_frequencies("Decade", 125000, 1.0, 1e8)
    lanes = 65536 if quick else 262144
    sel = torch.from_numpy(allf[:lanes].copy()).cuda()
    outbuf = torch.empty((lanes, c.n), dtype=torch.complex128, device="cuda")
    ms = timed(lambda: plan.solve_ac(sel, out=outbuf), reps=3, warm=1)
    alg = int(plan.info.alg_bytes_per_point)
    out.append({"config": 3, "what": f"RC ladder 5000 AC, first {lanes} of {len(allf)} decade-grid frequencies (n=5001, c64)", "ms": ms,
                "points_per_s": lanes / (ms * 1e-3), "kernel_ms_last_chunk": kb.last_kernel_ms(), "executor": plan.info.executor,
                "alg_bytes_per_point": alg, "model_hbm_bytes_per_point": int(plan.info.hbm_bytes_per_point),
                "roofline_frac_alg": alg * lanes / (ms * 1e-3) / 1e9 / peak,
                "full_sweep_estimate_s": len(allf) / (lanes / (ms * 1e-3)), "total_frequencies": int(len(allf))})
    del outbuf, sel
    # config 4: dual_rc_ladder transient, 10^4 steps of 5 us, 16384 Monte-Carlo variants of R1 R2 C1 C2
    c = kb.parse_circuit_description(golden["netlists"]["dual_rc_ladder"])
    names = ["R1", "R2", "C1", "C2"]
    varied = [c.element_index(nm) for nm in names]
    lanes, nsteps = (4096, 1000) if quick else (16384, 10000)
    vals = np.array([[c.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(lanes)] for j, e in enumerate(varied)])
    vals_d = torch.from_numpy(np.ascontiguousarray(vals)).cuda()
    pdc = kb.Plan(c, api.MODE_DC, varied=varied)
    ptr = kb.Plan(c, api.MODE_TRAN, varied=varied, time_step=5e-6)
    outbuf = torch.empty((lanes, nsteps + 1, c.n), dtype=torch.float64, device="cuda")
    ms = timed(lambda: ptr.solve_transient(pdc, nsteps, lanes, varied_values=vals_d, out=outbuf), reps=2, warm=1)
    pts = lanes * nsteps
    ptrap = kb.Plan(c, api.MODE_TRAN, varied=varied, time_step=5e-6, integrator=api.INTEGRATOR_TRAP)
    ms_t = timed(lambda: ptrap.solve_transient(pdc, nsteps, lanes, varied_values=vals_d, out=outbuf), reps=2, warm=1)
    out.append({"config": 4, "what": f"dual_rc_ladder transient TRAPEZOIDAL, {nsteps} steps x {lanes} MC lanes (n=4, f64)", "ms": ms_t,
                "points_per_s": pts / (ms_t * 1e-3), "executor": ptrap.info.executor, "scratch_slots": int(ptrap.info.scratch_slots),
                "spill_planes": int(ptrap.info.spill_planes)})
    out.append({"config": 4, "what": f"dual_rc_ladder transient BE_KRETS, {nsteps} steps x {lanes} MC lanes (n=4, f64)", "ms": ms,
                "points_per_s": pts / (ms * 1e-3), "executor": ptr.info.executor, "scratch_slots": int(ptr.info.scratch_slots),
                "spill_planes": int(ptr.info.spill_planes), "kernel_ms": kb.last_kernel_ms(),
                "note": "steps by krb_dense_tran<4> (native, unknowns in registers) unless KRB_NO_DENSE_TRAN is set (lane VM)"})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
