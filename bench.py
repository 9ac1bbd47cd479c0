#!/usr/bin/env python
"""Headline benchmark: solved sweep points/s of the batched stamp -> sparse LU -> solve path (BASELINE.json).

Workload at every N: BASELINE config 2 — resistor_ladder_5000 (n = 5001, nnz = nnz_lu = 15 000), DC sweep of V1 over
65 536 values per GPU (weak scaling: rank r sweeps values [r*65536 + i] * 2^-16).  One "step" = the whole 65 536-lane
batch: every lane is stamped, refactorised and solved and all 5001 unknowns of every point are delivered point-major.

  value  : device-resident throughput (inputs and outputs in HBM), CUDA events, max over ranks
  e2e    : same metric through the C ABI with pinned HOST buffers (H2D of the sweep values, D2H of every result), beside the box's
           bare pinned-D2H ceiling measured in the same run (e2e.link_ceiling_GBps)
  roofline: the block-partitioned chain kernel (krb_chain_block_kernel, one CTA per lane) against the measured HBM peak.
           frac = the bytes that must cross HBM (P + s*n per point: inputs in, results out — the kernel keeps the factorisation
           on chip) / kernel time / peak; frac_alg_model = the same time against SURVEY.md §8(d)'s 3-pass contract bytes
           (exceeds 1 for a fused kernel: a model figure, not a bus fraction)
  parity_spot_check: 64 random lanes of the timed output against the CPU oracle (+ linearity), computed after the timed region
  other_configs: BASELINE configs 3, 4 (trapezoidal and reference backward Euler), 5 (mesh AC / transient) and the ladder with all
           5000 resistors per-lane, measured in the same run under the same rules, each with its own roofline block
  cpu_baseline / --impl reference: the CPU oracle (restated krets semantics; the krets binary cannot be built here)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

LADDER_N = 5000
LANES_PER_GPU = 65536
STEP = 2.0 ** -16
FP64_PEAK_TFLOPS = 37.2   # measured, see roofline.compute.peak_source
METRIC = "solved sweep points/sec (batched sparse LU)"
UNIT = "points/s"
ALL_WORKLOADS = ["ladder_dc", "rc_ac", "dual_rc_tran", "dual_rc_tran_be", "mesh_tran", "mesh_ac", "ladder_mc", "rectifier_tran", "diode_dc_seq"]
OTHER_WORKLOADS = ["rc_ac", "ladder_mc", "dual_rc_tran", "dual_rc_tran_be", "mesh_ac", "mesh_tran", "rectifier_tran", "diode_dc_seq"]   # other_configs of the default line


def workload_config(n_gpus):
    return {
        "workload": "circuits/resistor_ladder_5000 DC sweep of V1 over 65,536 points per GPU (real f64 batched LU)",
        "n_unknowns": LADDER_N + 1, "nnz": 3 * LADDER_N, "lanes_per_gpu": LANES_PER_GPU, "lanes_total": LANES_PER_GPU * n_gpus,
        "sharding": f"independent lanes, {n_gpus} rank(s), no data-path collective",
        "l2": "per-step working set (2.6 GB of results per step, inputs 0.5 MB) >> 126 MB L2; no explicit flush",
        "result_layout": "point-major x[lane][unknown] (one vector per solved point, as the reference returns)",
    }


class ClockSampler:
    """SM clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled every ~2 ms from a thread
    (the timed region of the default run is only tens of ms); nvidia-smi is the fallback when pynvml is unavailable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self._stop, self._t, self.source = index, [], [], set(), threading.Event(), None, None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: resolve through the PCI bus id of the torch device
            import torch
            bus = torch.cuda.get_device_properties(index).pci_bus_id if hasattr(torch.cuda.get_device_properties(index), "pci_bus_id") else None
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index) if bus is None else None
            if self._h is None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if pynvml.nvmlDeviceGetPciInfo(h).bus == bus:
                        self._h = h
                        break
                if self._h is None:
                    self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nv = pynvml
            self.source = "nvml"
        except Exception:
            self._h = None
            self.source = "nvidia-smi"

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    nv = self._nv
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                    self._stop.wait(0.002)
                    continue
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    r = [c.strip() for c in out.split(",")]
                    self.sm.append(float(r[0])); self.mx.append(float(r[1]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                        if v.lower().startswith("active"):
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(sm), "source": self.source}


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank's host threads to the CPUs NVML reports as local to its GPU, so that the pinned result buffers it allocates
    (first touch) sit on that NUMA node: with several ranks the end-to-end path is bound by host-memory / PCIe-root traffic, and
    remote-node buffers halve it.  Best effort: any failure leaves the default placement."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_traffic(key):
    """dram bytes per launch of the named kernel from the committed ncu capture (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return json.load(f).get(key)
        except Exception:
            return None
    return None


def chain_block_shape(plan):
    """(T, M) of the block-partitioned chain kernel when the plan is admitted to it, else None (krb_plan_chain_block_info)."""
    import ctypes as C
    from krets_b200 import _lib
    L_ = _lib.lib()
    L_.krb_plan_chain_block_info.restype = C.c_int32
    L_.krb_plan_chain_block_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    info, nd = (C.c_int32 * 4)(), C.c_double(0.0)
    if L_.krb_plan_chain_block_info(plan._h, info, C.byref(nd)) != 0 or not info[0] or os.environ.get("KRB_CHAIN_SEQ"):
        return None
    return int(info[1]), int(info[2])


def cpu_baseline_sample(threads, points):
    from oracle import oracle
    import netgen
    oc = oracle.Circuit(netgen.resistor_ladder(LADDER_N))
    t0 = time.perf_counter()
    oc.solve_dc("V1", 0.0, (points - 1) * STEP, STEP, nthreads=threads)
    return points / (time.perf_counter() - t0)


def cpu_best_effort_sample(threads, points):
    """SURVEY 8(d) cpu_best_effort: analysis once, per point re-stamp + numeric LU + solve, OpenMP over points."""
    from oracle import oracle
    import netgen
    oc = oracle.Circuit(netgen.resistor_ladder(LADDER_N))
    t0 = time.perf_counter()
    oc.solve_dc_best_effort("V1", 0.0, (points - 1) * STEP, STEP, nthreads=threads)
    return points / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's CPU path (restated: oracle port) on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    import netgen
    oracle.build()
    threads = os.cpu_count() or 1
    sample = 256 * threads
    oc = oracle.Circuit(netgen.resistor_ladder(LADDER_N))
    for _ in range(args.warmup):
        oc.solve_dc("V1", 0.0, (sample - 1) * STEP, STEP, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oc.solve_dc("V1", 0.0, (sample - 1) * STEP, STEP, nthreads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} consecutive sweep points of the same workload per step (of 65,536), all {threads} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                         "note": "restated krets semantics (oracle/), not the krets binary: no Rust toolchain in this image; per point: "
                                 "re-stamp, hash-sum triplets, CSC build, symbolic+numeric sparse LU, solve"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))



# ---- workloads (BASELINE.json configs 2-5 + the per-lane stamping workload VERDICT r1 asked for) -----------------------------
def mc_matrix(circuit, varied, lanes, lane0):
    """values[k][lane] = nominal * (0.95 + 0.10 u), u = (splitmix64(seed + n_elements*lane + element) >> 11) * 2^-53; lane 0 nominal
    (same rule as tests/netgen.mc_factor, vectorised)."""
    import numpy as np
    lane = np.arange(lane0, lane0 + lanes, dtype=np.uint64)[None, :]
    e = np.asarray(varied, dtype=np.uint64)[:, None]
    with np.errstate(over="ignore"):
        x = np.uint64(0x6B72657473) + np.uint64(circuit.num_elements) * lane + e + np.uint64(0x9E3779B97F4A7C15)
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    f = 0.95 + 0.10 * ((z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53)
    f[:, (lane[0] == 0)] = 1.0
    nominal = np.array([circuit.element_value(int(k)) for k in varied])[:, None]
    return nominal * f


class Workload:
    """One bench workload: `step()` = one pass of the hot path over one batch resident in HBM.
    bytes: `alg` = SURVEY 8(d) contract figure per point; `moved` = bytes that must cross HBM per point for this executor
    (inputs + results for the fused kernels; the factor-store model for the front executor) — roofline.frac is on `moved`."""
    def __init__(self, **kw):
        self.extra = {}
        self.kernel_is_step = False      # True: roofline time = whole step (several kernels), else the lane kernel's own duration
        self.__dict__.update(kw)


def make_workload(name, args, rank):
    import numpy as np
    import torch
    import netgen
    import krets_b200 as kb
    from krets_b200 import api

    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        golden = json.load(f)
    dflt = args.lanes == LANES_PER_GPU
    if name == "ladder_dc":          # BASELINE config 2: the headline
        circuit = kb.parse_circuit_description(netgen.resistor_ladder(LADDER_N))
        plan = kb.Plan(circuit, api.MODE_DC, sweep_element=circuit.element_index("V1"), force_lane_vm=args.force_vm)
        lanes, n = args.lanes, circuit.n
        vals = (torch.arange(lanes, dtype=torch.float64, device="cuda") + float(rank * lanes)) * STEP
        out = torch.empty((lanes, n), dtype=torch.float64, device="cuda")
        shape = chain_block_shape(plan)
        kname = ("krb_lane_vm<f64>" if plan.info.executor != 1 else f"krb_chain_block_kernel<f64, {shape[0]}, {shape[1]}, SD>" if shape else "krb_chain_kernel<f64>")

        def check(k=64):
            from oracle import oracle
            oc = oracle.Circuit(netgen.resistor_ladder(LADDER_N))
            idx = np.random.default_rng(2).choice(lanes, size=min(k, lanes), replace=False)
            host, v = out[torch.from_numpy(idx).cuda()].cpu().numpy(), vals.cpu().numpy()
            return {"lanes_checked": int(len(idx)), "max_rel_err_vs_oracle": max(_rel_err(host[i], oc.solve_dc("V1", float(v[l]), float(v[l]), 1.0)[0]) for i, l in enumerate(idx))}
        return Workload(name=name, config=2, plan=plan, circuit=circuit, lanes=lanes, points=lanes, dtype="f64", s=8, out=out,
                        step=lambda: plan.solve_dc_sweep(vals, out=out), vals=vals, kernel=kname + " (stamp + LDU refactor + forward/back substitution fused per lane)",
                        what=workload_config(1)["workload"], alg=int(plan.info.alg_bytes_per_point), moved=8 + 8 * n if shape else int(plan.info.hbm_bytes_per_point),
                        moved_what="P + s*n: the sweep value in, n results out (SURVEY 8(d) fused floor B_floor)" if shape else "plan model: spilled factor planes + results",
                        traffic_key="krb_chain_block_kernel_dc_config2_bytes_per_launch" if shape and dflt else None, check=check)
    if name == "ladder_mc":          # north-star (a): per-lane stamping of device values at size — every resistor is a per-lane input
        circuit = kb.parse_circuit_description(netgen.resistor_ladder(LADDER_N))
        varied = [circuit.element_index(f"R{k}") for k in range(1, LADDER_N + 1)]
        plan = kb.Plan(circuit, api.MODE_DC, varied=varied)
        lanes, n = (32768 if dflt else args.lanes), circuit.n
        host_vals = mc_matrix(circuit, varied, lanes, rank * lanes)
        vals = torch.from_numpy(host_vals).cuda()
        out = torch.empty((lanes, n), dtype=torch.float64, device="cuda")
        shape = chain_block_shape(plan)

        def check(k=4):
            from oracle import oracle
            idx = np.random.default_rng(3).choice(lanes, size=min(k, lanes), replace=False)
            host, worst = out[torch.from_numpy(idx).cuda()].cpu().numpy(), 0.0
            for i, l in enumerate(idx):
                oc = oracle.Circuit(netgen.resistor_ladder(LADDER_N))
                for kk, e in enumerate(varied):
                    oc.set_element_value(e, host_vals[kk, l])
                worst = max(worst, _rel_err(host[i], oc.solve_op()))
            return {"lanes_checked": int(len(idx)), "max_rel_err_vs_oracle": worst}
        return Workload(name=name, config="2-mc", plan=plan, circuit=circuit, lanes=lanes, points=lanes, dtype="f64", s=8, out=out,
                        step=lambda: plan.solve_op(lanes, varied_values=vals, out=out),
                        kernel=(f"krb_chain_block_kernel<f64, {shape[0]}, {shape[1]}, !SD> (every row's stamps decoded from per-lane inputs)" if shape else "krb_chain_kernel<f64>"),
                        what=f"resistor_ladder_5000 operating point, {lanes} Monte-Carlo lanes per GPU with ALL 5000 resistors per-lane (+-5 %): 40 KB of element values stamped per lane",
                        alg=int(plan.info.alg_bytes_per_point), moved=8 * len(varied) + 8 * n,
                        moved_what="P + s*n: 5000 per-lane resistor values in, n results out", traffic_key=None, check=check)
    if name == "rc_ac":              # BASELINE config 3
        ladder_n = int(os.environ.get("KRB_BENCH_LADDER_N", LADDER_N))      # profiling: other ladder lengths
        circuit = kb.parse_circuit_description(netgen.rc_ladder(ladder_n))
        plan = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1.0)
        allf = kb.generate_frequencies("Decade", 125000, 1.0, 1e8)
        lanes = 131072 if dflt else args.lanes
        sel_host = allf[(rank * lanes) % len(allf):][:lanes].copy()
        sel = torch.from_numpy(sel_host).cuda()
        lanes, n = int(sel.shape[0]), circuit.n
        out = torch.empty((lanes, n), dtype=torch.complex128, device="cuda")
        shape = chain_block_shape(plan)

        def check(k=16):
            from oracle import oracle
            oc = oracle.Circuit(netgen.rc_ladder(ladder_n))
            idx = np.sort(np.random.default_rng(4).choice(lanes, size=min(k, lanes), replace=False))
            host = out[torch.from_numpy(idx).cuda()].cpu().numpy()
            _, ref = oc.solve_ac(sel_host[idx], nthreads=min(os.cpu_count() or 1, 8))
            return {"lanes_checked": int(len(idx)), "max_rel_err_vs_oracle": max(_rel_err(host[i], ref[i]) for i in range(len(idx)))}
        return Workload(name=name, config=3, plan=plan, circuit=circuit, lanes=lanes, points=lanes, dtype="c64", s=16, out=out,
                        step=lambda: plan.solve_ac(sel, out=out),
                        kernel=(f"krb_chain_block_kernel<c64, {shape[0]}, {shape[1]}, SD>" if shape else "krb_chain_kernel<c64>"),
                        what=f"resistor_ladder_5000-class RC ladder AC sweep, {lanes} of {len(allf)} log-spaced frequencies per GPU (complex f64); the full sweep is {len(allf)} lanes = {len(allf) * n * 16 / 1e9:.0f} GB of results",
                        alg=int(plan.info.alg_bytes_per_point), moved=8 + 16 * n if shape else int(plan.info.hbm_bytes_per_point),
                        moved_what="P + s*n: the frequency in, n complex results out",
                        traffic_key="krb_chain_block_kernel_ac_config3_131072_lanes_bytes_per_launch" if shape and dflt and ladder_n == LADDER_N else None, check=check)
    if name in ("dual_rc_tran", "dual_rc_tran_be"):     # BASELINE config 4: trapezoidal (as written) and the reference's backward Euler
        integ = api.INTEGRATOR_TRAP if name == "dual_rc_tran" else api.INTEGRATOR_BE_KRETS
        text = golden["netlists"]["dual_rc_ladder"]
        circuit = kb.parse_circuit_description(text)
        names = ["R1", "R2", "C1", "C2"]
        varied = [circuit.element_index(nm) for nm in names]
        lanes = 16384 if dflt else args.lanes
        nsteps = args.tran_steps or 10000
        host_vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 4) for l in range(rank * lanes, (rank + 1) * lanes)] for j, e in enumerate(varied)])
        vals = torch.from_numpy(host_vals).cuda()
        pdc = kb.Plan(circuit, api.MODE_DC, varied=varied)
        plan = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=5e-6, integrator=integ)
        out = torch.empty((lanes, nsteps + 1, circuit.n), dtype=torch.float64, device="cuda")

        def check(k=2):
            from oracle import oracle
            idx = np.random.default_rng(5).choice(lanes, size=min(k, lanes), replace=False)
            worst = 0.0
            for l in idx:
                oc = oracle.Circuit(text)
                for j, nm in enumerate(names):
                    oc.set_element_value(oc.element_index(nm), host_vals[j, l])
                _, ref = oc.solve_tran(5e-6, nsteps * 5e-6, integrator=integ)
                worst = max(worst, _rel_err(out[int(l)].cpu().numpy(), ref))
            return {"lanes_checked": int(len(idx)), "rows_per_lane": nsteps + 1, "max_rel_err_vs_oracle": worst}
        return Workload(name=name, config=4, plan=plan, circuit=circuit, lanes=lanes, points=lanes * nsteps, dtype="f64", s=8, out=out,
                        step=lambda: plan.solve_transient(pdc, nsteps, lanes, varied_values=vals, out=out),
                        kernel="krb_dense_tran<4> (per-lane propagator steps, factors on chip)",
                        what=f"circuits/dual_rc_ladder transient, {nsteps} {'trapezoidal' if integ == api.INTEGRATOR_TRAP else 'backward-Euler (reference: transient.rs:43-112, incl. its source lag)'} steps x {lanes} Monte-Carlo R/C variants per GPU",
                        alg=8 * (int(plan.info.nnz_lu) + 3 * circuit.n), alg_what="B_step = s*(nnz_lu + 3*n), SURVEY 8(d)",
                        moved=8 * circuit.n, moved_what="s*n: one result row per solved point (factors and state stay on chip)",
                        traffic_key="krb_dense_tran_config4_bytes_per_launch" if dflt and not args.tran_steps else None, check=check)
    if name == "rectifier_tran":
        # The nonlinear path (north star (d)): the reference's own rectifier test circuit (circuits/rectifier: four diodes, R, C, two sine
        # sources; test_transient.rs) stepped with backward Euler, Newton-Raphson with the per-lane convergence mask at every step, on the
        # lane VM (specialised kernel, vm_jit.cpp); Monte-Carlo lanes vary R1 / C1.  Not a BASELINE config: it puts a number on the nonlinear path.
        text = golden["netlists"]["rectifier"]
        circuit = kb.parse_circuit_description(text)
        names = ["C1", "R1"]
        varied = [circuit.element_index(nm) for nm in names]
        lanes = 16384 if dflt else args.lanes
        nsteps = args.tran_steps or 1000
        h = 50e-6
        host_vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 2) for l in range(rank * lanes, (rank + 1) * lanes)] for j, e in enumerate(varied)])
        vals = torch.from_numpy(host_vals).cuda()
        pdc = kb.Plan(circuit, api.MODE_DC, varied=varied)
        plan = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=h)
        out = torch.empty((lanes, nsteps + 1, circuit.n), dtype=torch.float64, device="cuda")

        def check(k=2):
            from oracle import oracle
            idx = np.random.default_rng(7).choice(lanes, size=min(k, lanes), replace=False)
            worst = 0.0
            for l in idx:
                oc = oracle.Circuit(text)
                for j, nm in enumerate(names):
                    oc.set_element_value(oc.element_index(nm), host_vals[j, l])
                _, ref = oc.solve_tran(h, nsteps * h)
                worst = max(worst, _rel_err(out[int(l)].cpu().numpy(), ref))
            return {"lanes_checked": int(len(idx)), "rows_per_lane": nsteps + 1, "max_rel_err_vs_oracle": worst,
                    "tolerance": "1e-6: Newton-Raphson with exp(); libm and CUDA exp differ in the last ulp (tests/test_gpu_parity.py::test_transient_rectifier)"}
        return Workload(name=name, config="nonlinear", plan=plan, circuit=circuit, lanes=lanes, points=lanes * nsteps, dtype="f64", s=8, out=out,
                        step=lambda: plan.solve_transient(pdc, nsteps, lanes, varied_values=vals, out=out),
                        kernel="krb_jit_vm (the lane program as a circuit-specialised kernel, NVRTC for sm_100a at first use: stamp + LDU + solve per Newton iteration, per-lane convergence mask; KRB_VM_JIT=0: krb_lane_vm interpreter)",
                        what=f"circuits/rectifier (4 diodes) transient, {nsteps} backward-Euler steps x {lanes} Monte-Carlo R/C variants per GPU, Newton-Raphson at every step (nonlinear path; not a BASELINE config)",
                        alg=8 * (int(plan.info.nnz_lu) + 3 * circuit.n), alg_what="B_step = s*(nnz_lu + 3*n) per Newton iteration (at least two per point), SURVEY 8(d)",
                        moved=8 * circuit.n, moved_what="s*n: one result row per solved point (the lane program's working set stays on chip)",
                        traffic_key=None, check=check)
    if name == "diode_dc_seq":
        # The reference's nonlinear DC sweep as written (dc.rs:47-113): the sweep points of circuits/diode_iv_curve are solved in order,
        # Newton-Raphson warm-started from the previous point (dc.rs:58,112) — the points are the rows of a lane, the lanes are Monte-Carlo
        # variants of R1.  Not a BASELINE config: it puts a number on krb_solve_dc_sweep_sequential.
        text = golden["netlists"]["diode_iv_curve"]
        circuit = kb.parse_circuit_description(text)
        varied = [circuit.element_index("R1")]
        lanes = 65536 if dflt else args.lanes
        npts, start, step_size = 101, 0.0, 0.02
        host_vals = np.array([[circuit.element_value(e) * netgen.mc_factor(l, j, 2) for l in range(rank * lanes, (rank + 1) * lanes)] for j, e in enumerate(varied)])
        vals = torch.from_numpy(host_vals).cuda()
        plan = kb.Plan(circuit, api.MODE_DC, sweep_element=circuit.element_index("V1"), varied=varied)
        out = torch.empty((lanes, npts, circuit.n), dtype=torch.float64, device="cuda")

        def check(k=2):
            from oracle import oracle
            idx = np.random.default_rng(7).choice(lanes, size=min(k, lanes), replace=False)
            worst = 0.0
            for l in idx:
                oc = oracle.Circuit(text)
                oc.set_element_value(oc.element_index("R1"), host_vals[0, l])
                ref = oc.solve_dc("V1", start, start + (npts - 1) * step_size, step_size, warm_start=True)
                worst = max(worst, _rel_err(out[int(l)].cpu().numpy(), ref[:npts]))
            return {"lanes_checked": int(len(idx)), "rows_per_lane": npts, "max_rel_err_vs_oracle": worst,
                    "oracle": "the reference's sequential warm-started sweep (dc.rs:58,112)"}
        return Workload(name=name, config="nonlinear-dc", plan=plan, circuit=circuit, lanes=lanes, points=lanes * npts, dtype="f64", s=8, out=out,
                        step=lambda: plan.solve_dc_sweep_sequential(start, step_size, npts, lanes, varied_values=vals, out=out),
                        kernel="krb_jit_vm, one launch per sweep point (rows of a lane in order, warm-started from the previous row; KRB_VM_JIT=0: krb_lane_vm)",
                        what=f"circuits/diode_iv_curve DC sweep of V1, {npts} points in order (Newton-Raphson warm-started from the previous point, dc.rs:58) x {lanes} Monte-Carlo R1 variants per GPU (nonlinear path; not a BASELINE config)",
                        alg=8 * (int(plan.info.nnz_lu) + 3 * circuit.n), alg_what="s*(nnz_lu + 3*n) per Newton iteration (two per point), SURVEY 8(d)",
                        moved=8 * circuit.n, moved_what="s*n: one result row per solved point (the lane program's working set stays on chip)",
                        traffic_key=None, check=check)
    if name in ("mesh_ac", "mesh_tran"):                # BASELINE config 5
        transient = name == "mesh_tran"
        g = args.mesh_grid
        text = netgen.rlc_mesh(g, g, transient=transient)
        circuit = kb.parse_circuit_description(text)
        kinds = [circuit.element(e).kind for e in range(circuit.num_elements)]
        varied = [e for e, k in enumerate(kinds) if k in (api.KIND_R, api.KIND_C, api.KIND_L)]
        # lanes per GPU: whole waves of the one-CTA-per-SM launches of the top tree levels (a single front per level: lanes CTAs)
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        lanes = ((4 if transient else 2) * sms) if dflt else args.lanes
        host_vals = mc_matrix(circuit, varied, lanes, rank * lanes)
        vals = torch.from_numpy(host_vals).cuda()
        n = circuit.n
        if transient:
            nsteps = args.tran_steps or 32
            pdc = kb.Plan(circuit, api.MODE_DC, varied=varied, executor=api.EXEC_STAGED)
            plan = kb.Plan(circuit, api.MODE_TRAN, varied=varied, time_step=1e-9, integrator=api.INTEGRATOR_TRAP, executor=api.EXEC_STAGED)
            out = torch.empty((lanes, nsteps + 1, n), dtype=torch.float64, device="cuda")
            step = lambda: plan.solve_transient(pdc, nsteps, lanes, varied_values=vals, out=out)
            points, dtype, s = lanes * (nsteps + 1), "f64", 8
            what = f"synthetic {g}x{g} RLC mesh (n={n}) transient Monte-Carlo: {lanes} lanes per GPU x ({nsteps} trapezoidal steps + operating point)"
            alg, alg_what = s * (int(plan.info.nnz_lu) + 3 * n), "B_step = s*(nnz_lu + 3*n) per solved time point, SURVEY 8(d)"

            def check(k=1):
                return _mesh_check(text, varied, host_vals, out, 0, "tran", 1e-9, nsteps)
        else:
            plan = kb.Plan(circuit, api.MODE_AC, varied=varied, nominal_frequency=1e4, executor=api.EXEC_STAGED)
            freqs = torch.logspace(4, 8, lanes, dtype=torch.float64, device="cuda")
            out = torch.empty((lanes, n), dtype=torch.complex128, device="cuda")
            step = lambda: plan.solve_ac(freqs, varied_values=vals, out=out)
            points, dtype, s = lanes, "c64", 16
            what = f"synthetic {g}x{g} RLC mesh (n={n}) AC Monte-Carlo: {lanes} lanes per GPU, 10 kHz..100 MHz"
            alg, alg_what = int(plan.info.alg_bytes_per_point), "P + s*(4*nnz_lu + 3*n), SURVEY 8(d)"

            def check(k=1):
                return _mesh_check(text, varied, host_vals, out, lanes // 2, "ac", float(freqs[lanes // 2].item()), 0)
        w = Workload(name=name, config=5, plan=plan, circuit=circuit, lanes=lanes, points=points, dtype=dtype, s=s, out=out, step=step,
                     kernel="krb_front_factor / krb_front_forward / krb_front_backward (supernodal multifrontal LU, level-scheduled)",
                     what=what, alg=alg, alg_what=alg_what, moved=alg,
                     moved_what="the contract bytes themselves: the factor store does not fit on chip, so stamp -> factor -> solve streams it through HBM",
                     traffic_key=None, check=check)
        w.kernel_is_step = True
        # FP64 work of one step (B200's FP64 pipe, not HBM, is the nearer roof of the supernodal factorisation): the plan compiler's count
        # of real flops, factor + one pair of triangular solves per plan (AC: the two refinement solves and residuals are not counted)
        if transient:
            w.flops_per_step = lanes * (int(pdc.info.flops_per_point) + int(plan.info.flops_per_point) + (nsteps - 1) * 2 * int(plan.info.nnz_lu))
            w.flops_what = "lanes x (DC-plan factor + solve, transient-plan factor + solve, 2*nnz_lu per further step); krb_plan_info.flops_per_point"
        else:
            w.flops_per_step = lanes * int(plan.info.flops_per_point)
            w.flops_what = "lanes x (complex factor + one solve, 4 real multiply-adds per complex one); krb_plan_info.flops_per_point"
        w.extra["nnz_lu"] = int(plan.info.nnz_lu)
        w.heavy = True
        return w
    raise ValueError(name)


def _rel_err(a, b):
    """tests/helpers.rel_err: max |a-b| / max(|b|, 1e-6 ||b||_inf) (north-star tolerance 1e-9)."""
    import numpy as np
    a, b = np.asarray(a), np.asarray(b)
    top = float(np.max(np.abs(b))) if b.size else 0.0
    scale = np.maximum(np.abs(b), 1e-6 * top if top > 0 else 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if b.size else 0.0


def _mesh_check(text, varied, host_vals, out, lane, mode, w, nsteps):
    """config 5 is far beyond the oracle's natural-order LU: residual on the ORACLE-stamped matrix + SuperLU on the same matrix
    (tests/mesh_reference.py), one lane."""
    import numpy as np
    import scipy.sparse.linalg as spl
    import mesh_reference as mr
    from oracle import oracle
    oc = oracle.Circuit(text)
    tab = mr.ElementTable(oc)
    for k, e in enumerate(varied):
        oc.set_element_value(e, host_vals[k, lane])
    x = out[lane].cpu().numpy()
    if mode == "ac":
        A, b = mr.csc(oc, "ac", w), mr.ac_rhs(oc, tab, tab.values(oc))
        return {"lanes_checked": 1, "residual_on_oracle_matrix": float(np.max(np.abs(A @ x - b)) / np.max(np.abs(b))),
                "max_rel_err_vs_superlu_on_oracle_matrix": _rel_err(x, spl.splu(A).solve(b))}
    rows = min(nsteps, 4)
    ref = mr.trap_rows(oc, tab, w, rows)
    return {"lanes_checked": 1, "rows_checked": rows + 1, "max_rel_err_vs_superlu_stepping_on_oracle_matrices": _rel_err(x[: rows + 1], ref)}


def time_workload(w, W, K, ctx):
    """W warm-up steps, then exactly K timed steps between barrier + synchronize, CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    import krets_b200 as kb
    for _ in range(W):
        _, status = w.step()
    ctx["barrier"]()
    launches0 = kb.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(ctx["local_rank"]) as clocks:
        ctx["barrier"]()
        ev0.record()
        for _ in range(K):
            w.step()
        ev1.record()
        ctx["barrier"]()
    launches = kb.kernel_launch_count() - launches0
    ms = ev0.elapsed_time(ev1)
    # lane-kernel duration, measured live with CUDA events on the launch stream inside the library (separate pass so that the
    # event reads do not perturb the timed loop above)
    kernel_ms = 0.0
    reps = 1 if getattr(w, "heavy", False) else K
    for _ in range(reps):
        w.step()
        kernel_ms += kb.last_kernel_ms()
    kernel_ms /= reps
    t = torch.tensor([ms, kernel_ms], dtype=torch.float64, device="cuda")
    if ctx["distributed"]:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, kernel_ms = float(t[0].item()), float(t[1].item())
    world = ctx["world"]
    peak, peak_src = measured_peaks()
    t_roof = (ms_max / K if w.kernel_is_step else kernel_ms) * 1e-3
    moved_GBps = w.moved * w.points / t_roof / 1e9
    alg_GBps = w.alg * w.points / t_roof / 1e9
    traffic = ncu_traffic(w.traffic_key) if w.traffic_key else None
    res = {
        "workload": w.what, "baseline_config": w.config, "dtype": w.dtype, "n_unknowns": w.circuit.n, "lanes_per_gpu": w.lanes,
        "value": w.points * world * K / (ms_max * 1e-3), "unit": UNIT, "steps": K, "warmup": W, "ms_per_step": ms_max / K, "gpu_launches": int(launches),
        "lanes_with_status": int((status != 0).sum().item()),
        "roofline": {"bound": "hbm", "achieved": moved_GBps, "peak": peak, "unit": "GB/s", "frac": moved_GBps / peak, "traffic": traffic,
                     "kernel": w.kernel, "kernel_ms": kernel_ms, "timed": "whole step (all kernels of the step)" if w.kernel_is_step else "the lane kernel alone, CUDA events on its stream",
                     "bytes_per_point": w.moved, "bytes_model": w.moved_what, "peak_source": peak_src,
                     "frac_of_nominal_8TBps": moved_GBps / 8000.0,
                     "alg_bytes_per_point": w.alg, "alg_bytes_model": getattr(w, "alg_what", "P + s*(4*nnz_lu + 3*n), SURVEY 8(d)"),
                     "frac_alg_model": alg_GBps / peak,
                     "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, round-2 capture (profiles/traffic.json)" if traffic else None},
        **w.extra,
    }
    if getattr(w, "flops_per_step", None):
        tf = w.flops_per_step / t_roof / 1e12
        res["roofline"]["compute"] = {"bound": "fp64", "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf / FP64_PEAK_TFLOPS,
                                      "flops_per_step": int(w.flops_per_step), "flops_model": w.flops_what,
                                      "peak_source": "tools/microbench/dmma_rate.cu on this pool's B200: DFMA and DMMA.8x8x4 both sustain 64 FMA lanes / clk / SM = 37.2 TFLOP/s at 1965 MHz (profiles/r2_notes.md)"}
    return res, clocks.summary()


def strong_scaling_rc_ac(ctx):
    """BASELINE config 3 as written: the WHOLE 10^6-frequency sweep of the RC ladder 5000, split over the ranks (strong scaling),
    solved in 65 536-lane chunks.  `value`: results stay sharded (the default); `with_gather`: every chunk travels to rank 0 while the
    next one is solved, gather inside the timed region."""
    import torch
    import torch.distributed as dist
    import netgen
    import krets_b200 as kb
    from krets_b200 import api
    from krets_b200.sharding import solve_sharded_chunked, shard_bounds
    world, rank = ctx["world"], ctx["rank"]
    circuit = kb.parse_circuit_description(netgen.rc_ladder(LADDER_N))
    plan = kb.Plan(circuit, api.MODE_AC, nominal_frequency=1.0)
    allf = torch.from_numpy(kb.generate_frequencies("Decade", 125000, 1.0, 1e8)).cuda()
    total, n, chunk = int(allf.shape[0]), circuit.n, 65536
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")

    def solve_chunk(lo, hi, dst):
        _, st = plan.solve_ac(allf[lo:hi], out=dst)
        bad.add_((st != 0).sum())

    def timed(gather_to, reps):
        solve_sharded_chunked(total, solve_chunk, chunk_lanes=chunk, row_shape=(n,), dtype=torch.complex128, device="cuda", gather_to=gather_to)
        ctx["barrier"]()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = None
        for _ in range(reps):
            res = None      # the previous sweep's results go back to the caching allocator first: no cudaMalloc of a second 80/N GB buffer inside the timed region
            lo, res = solve_sharded_chunked(total, solve_chunk, chunk_lanes=chunk, row_shape=(n,), dtype=torch.complex128, device="cuda", gather_to=gather_to)
        e1.record()
        ctx["barrier"]()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        if ctx["distributed"]:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        v1 = None
        if res is not None:
            col = circuit.index_map["V(1)"]
            v1 = float((res[:, col] - 1.0).abs().max().item())        # source row: V(1) = 1 on every lane
        del res
        torch.cuda.empty_cache()
        return float(t.item()), v1
    ms_sh, v1 = timed(None, 2)
    out = {"name": "rc_ac_full_strong", "baseline_config": 3, "scaling": "strong", "dtype": "c64", "n_unknowns": n, "lanes_total": total,
           "workload": f"resistor_ladder_5000-class RC ladder AC sweep, ALL {total} log-spaced frequencies split over {world} GPU(s) (strong scaling), {chunk}-lane chunks; {total * n * 16 / 1e9:.0f} GB of results",
           "value": total / (ms_sh * 1e-3), "unit": UNIT, "ms_per_sweep": ms_sh, "results": "stay sharded on their GPUs (default)",
           "max_abs_err_source_row": v1}
    if world > 1:
        ms_g, v1g = timed(0, 1)
        lo, hi = shard_bounds(total, world, 0)
        inbound = (total - (hi - lo)) * n * 16
        out["with_gather"] = {"value": total / (ms_g * 1e-3), "unit": UNIT, "ms_per_sweep": ms_g, "bytes_inbound_rank0": inbound, "GBps_inbound": inbound / (ms_g * 1e-3) / 1e9,
                              "collective": "chunked ncclSend/ncclRecv to rank 0, grouped per round, overlapped with the next chunk's solve, inside the timed region",
                              "max_abs_err_source_row": v1g}
    out["lanes_with_status"] = int(bad.item())
    return out if rank == 0 else None


def link_ceiling(ctx, nbytes=1 << 30, reps=3):
    """Bare pinned cudaMemcpyAsync D2H on every rank at once: the box's ceiling for delivering results to the host."""
    import torch
    import torch.distributed as dist
    src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst.copy_(src, non_blocking=True)
    ctx["barrier"]()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if ctx["distributed"]:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del src, dst
    return nbytes * reps * ctx["world"] / float(t.item()) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="krets_b200")
    ap.add_argument("--lanes", type=int, default=LANES_PER_GPU)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs block (BASELINE configs 3-5 + the Monte-Carlo ladder)")
    ap.add_argument("--force-vm", action="store_true", help="profiling: run the workload on the general lane VM instead of the chain executor")
    ap.add_argument("--workload", default="ladder_dc", choices=ALL_WORKLOADS,
                    help="ladder_dc = BASELINE config 2 (the headline, default); rc_ac = config 3, dual_rc_tran(_be) = config 4, mesh_* = config 5, "
                         "ladder_mc = ladder 5000 with every resistor per-lane")
    ap.add_argument("--mesh-grid", type=int, default=224)
    ap.add_argument("--tran-steps", type=int, default=0, help="time steps per lane for the transient workloads (0 = workload default)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import krets_b200 as kb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    cpus = bind_to_gpu_numa_node(local_rank) if distributed else None
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert kb.device_count() > 0, "bench needs a CUDA device (no CPU fallback)"
    W, K = max(args.warmup, 3), args.steps

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
    ctx = {"world": world, "rank": rank, "local_rank": local_rank, "distributed": distributed, "barrier": barrier}

    # ---- the main line: device-resident throughput of the selected workload ----
    w = make_workload(args.workload, args, rank)
    info = w.plan.info
    main_res, clocks = time_workload(w, W, K, ctx)
    lanes, n = w.lanes, w.circuit.n
    parity = w.check() if rank == 0 else None
    if w.name == "ladder_dc" and rank == 0:      # also the size-independent property: x(lane) = v_lane * x(1 V)
        ref = w.out[lanes - 1] / w.vals[lanes - 1]
        parity["linearity_max_rel_err"] = float(((w.out[lanes // 2] - ref * w.vals[lanes // 2]).abs().max() / ref.abs().max()).item())

    # ---- end to end through the C ABI with pinned host buffers (headline workload only) ----
    e2e = None
    if not args.no_e2e and w.name == "ladder_dc":
        out, vals, plan = w.out, w.vals, w.plan
        h_vals = torch.empty(lanes, dtype=torch.float64).pin_memory()
        h_vals.copy_(vals.cpu())
        h_out = torch.empty((lanes, n), dtype=torch.float64).pin_memory()
        hv, ho = h_vals.numpy(), h_out.numpy()
        for _ in range(2):
            plan.solve_dc_sweep(hv, out=ho)
        barrier()
        Ke = max(3, min(K, 10))
        t0 = time.perf_counter()
        for _ in range(Ke):
            plan.solve_dc_sweep(hv, out=ho)       # returns after the D2H of every result has completed
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d2h = lanes * n * 8 + lanes * 4
        e2e = {"value": lanes * world * Ke / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": lanes * 8, "d2h_bytes_per_step": d2h, "steps": Ke,
               "note": "krb_solve_dc_sweep with KRB_MEM_HOST pinned buffers; D2H of all 5001 unknowns of every point is inside the timed region",
               "host_cpus_rank0": (f"{cpus[0]}-{cpus[-1]} ({len(cpus)} CPUs local to the GPU, NVML)" if cpus else "default placement")}
        assert np.array_equal(ho[lanes - 1], out[lanes - 1].cpu().numpy())
        del h_out, ho
        e2e["d2h_GBps"] = d2h * world * Ke / float(t.item()) / 1e9
        e2e["link_ceiling_GBps"] = link_ceiling(ctx)
        e2e["frac_of_link_ceiling"] = e2e["d2h_GBps"] / e2e["link_ceiling_GBps"]
        e2e["link_ceiling_note"] = f"bare pinned cudaMemcpyAsync D2H of 1 GiB x 3 on all {world} rank(s) at once, aggregate"
        # EXTENSION (labelled; the reference returns every unknown, ac.rs:97-105): the same call with krb_io.select = 8 unknowns —
        # what a plotting front end reads.  Reported beside the full-delivery headline, never instead of it.
        sel = [0, 1, 2, n // 4, n // 2, (3 * n) // 4, n - 2, n - 1]
        h_sel = torch.empty((lanes, len(sel)), dtype=torch.float64).pin_memory()
        hs = h_sel.numpy()
        for _ in range(2):
            plan.solve_dc_sweep(hv, out=hs, select=sel)
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            plan.solve_dc_sweep(hv, out=hs, select=sel)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert np.array_equal(hs[lanes - 1], out[lanes - 1].cpu().numpy()[sel])
        e2e["selected"] = {"value": lanes * world * Ke / float(t.item()), "unit": UNIT, "n_select": len(sel), "d2h_bytes_per_step": lanes * len(sel) * 8 + lanes * 4,
                           "note": "EXTENSION, not reference behaviour: krb_io.select delivers 8 of the 5001 unknowns per point (host buffers, same call otherwise)"}

    # ---- results gathered on rank 0: by default they stay sharded (`value` above: no data-path collective); here every
    # chunk is sent to rank 0 over NVLink (ncclSend/ncclRecv, grouped per round) while the next chunk is being solved, and the
    # whole thing - solve + gather - is inside the timed region ----
    gather = None
    if distributed and w.name == "ladder_dc":
        from krets_b200.sharding import solve_sharded_chunked
        plan = w.plan
        total = lanes * world
        all_vals = torch.arange(total, dtype=torch.float64, device="cuda") * STEP
        chunk = 16384

        def solve_chunk(lo, hi, dst):
            plan.solve_dc_sweep(all_vals[lo:hi], out=dst)
        solve_sharded_chunked(total, solve_chunk, chunk_lanes=chunk, row_shape=(n,), dtype=torch.float64, device="cuda", gather_to=0)   # warm-up: NCCL connections
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        _, full = solve_sharded_chunked(total, solve_chunk, chunk_lanes=chunk, row_shape=(n,), dtype=torch.float64, device="cuda", gather_to=0)
        g1.record()
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        nbytes = lanes * n * 8 * (world - 1)
        if rank == 0:
            assert torch.equal(full[:lanes], w.out)
        gather = {"collective": f"chunked ncclSend/ncclRecv to rank 0 ({chunk}-lane chunks, grouped per round), overlapped with the solve of the next chunk; solve + gather timed together",
                  "ms": float(gms.item()), "bytes_inbound_rank0": nbytes, "GBps_inbound": nbytes / (float(gms.item()) * 1e-3) / 1e9,
                  "value_with_gather": total / (float(gms.item()) * 1e-3),
                  "note": "the root's inbound NVLink port (<= 900 GB/s), not the solve, bounds a gathered sweep: results stay sharded by default"}
        del full, all_vals

    # ---- the other BASELINE configs, measured in the same run under the same rules ----
    others = []
    if not args.no_others and w.name == "ladder_dc" and lanes == LANES_PER_GPU:
        del w
        torch.cuda.empty_cache()
        for name in OTHER_WORKLOADS:
            try:
                ow = make_workload(name, args, rank)
                heavy = getattr(ow, "heavy", False)
                res, _ = time_workload(ow, 1 if heavy else 3, 2 if heavy else max(3, min(K, 10)), ctx)
                if rank == 0:
                    res["parity_spot_check"] = ow.check()
                res["name"] = name
                others.append(res)
            except Exception as ex:   # a failing side workload must not lose the headline line
                others.append({"name": name, "error": f"{type(ex).__name__}: {ex}"})
            ow = None
            torch.cuda.empty_cache()

    if not args.no_others and args.workload == "ladder_dc" and lanes == LANES_PER_GPU:
        try:
            res = strong_scaling_rc_ac(ctx)
            if res is not None:
                others.append(res)
        except Exception as ex:
            others.append({"name": "rc_ac_full_strong", "error": f"{type(ex).__name__}: {ex}"})
        torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": main_res["dtype"], "data": "synthetic",
            "config": workload_config(world) if (args.workload == "ladder_dc" and lanes == LANES_PER_GPU) else
                      {"workload": main_res["workload"], "n_unknowns": n, "nnz": int(info.nnz), "lanes_per_gpu": lanes, "lanes_total": lanes * world,
                       "sharding": f"independent lanes, {world} rank(s), no data-path collective",
                       "l2": f"per-step working set ({main_res['roofline']['bytes_per_point'] * main_res['value'] * main_res['ms_per_step'] * 1e-3 / world / 1e9:.2f} GB per GPU and step on the roofline's byte model) >> 126 MB L2; no explicit flush",
                       "result_layout": "point-major x[lane][unknown] (one vector per solved point, as the reference returns)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": main_res["gpu_launches"], "roofline": main_res["roofline"],
            "plan": {"executor": {0: "lane_vm", 1: "chain", 2: "front"}[int(info.executor)], "nnz_lu": int(info.nnz_lu), "flops_per_point": int(info.flops_per_point)},
            "parity_spot_check": parity, "lanes_with_status": main_res["lanes_with_status"],
        }
        for k in ("nnz_lu",):
            if k in main_res:
                line["config"][k] = main_res[k]
        if gather:
            line["gather"] = gather
        if others:
            line["other_configs"] = others
        if world == 1 and not args.no_cpu and args.workload == "ladder_dc":
            from oracle import oracle
            oracle.build()
            threads = os.cpu_count() or 1
            v1 = cpu_baseline_sample(1, 2048)
            vN = cpu_baseline_sample(threads, 2048 * threads)
            vbe = cpu_best_effort_sample(threads, 4096 * threads)
            line["cpu_baseline"] = {"value": vN, "unit": UNIT, "cores": threads, "kind": "port",
                                    "best_effort_value": vbe,
                                    "best_effort_note": "cpu_best_effort of SURVEY 8(d): pattern, triplet->CSC slot map and buffers built once, per point re-stamp + "
                                                        f"numeric LU + solve, {threads} threads over {4096 * threads} points (bit-identical results to the faithful mode)",
                                    "sample": f"{2048 * threads} consecutive points of the same sweep (of 65,536), {threads} threads over points; "
                                              f"single-thread (krets is sequential): {v1:.1f} points/s on 2048 points",
                                    "single_thread_value": v1,
                                    "note": "oracle/ restatement of krets semantics, not the krets binary (no Rust toolchain in the image)"}
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
