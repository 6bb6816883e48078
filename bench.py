#!/usr/bin/env python
"""bench.py - dose+gradient evaluations per second on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2]

One step = one evaluation of the hot path: x = y*w_dose, d = D x, penalty (objective f and
voxel gradient), gb = D^T vg, aperture gradient g  (calcObjGrad, greedyVMATsampling.py:578-582
plus the beamlet gradient PPsubroutine :646 needs).  Workload at N=1: BASELINE.json
configs[1], the synthetic 300k-voxel x 10 080-beamlet x 180-control-point case at 1 % density.
For N>1 the same case is voxel-sharded (strong scaling) with one all-reduce per evaluation.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi SM clocks and throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


DEVICE_WORKLOADS = ("cfg3", "cfg4")      # too large to build through scipy on the host


def workload(name, scale, device=None):
    from vmatwpencode_b200.synth import make_config, make_device_config
    if name in DEVICE_WORKLOADS and device is not None:
        case = make_device_config(name, scale=scale, device=device)
    else:
        case = make_config(name, scale=scale)
    # all control points selected with a centred open aperture; intensities in (0.5, 1.5)
    w_dose = np.zeros(case.B)
    w_grad = np.zeros(case.B)
    for i in range(case.nb):
        lo = int(case.beam_ptr[i])
        for m in range(case.M):
            w_dose[lo + m * case.N + 1: lo + (m + 1) * case.N - 1] = 1.0
            w_grad[lo + m * case.N + 1: lo + (m + 1) * case.N - 1] = 1.0
    y = 0.5 + (np.arange(case.nb) * 0.6180339887) % 1.0
    return case, y, w_dose, w_grad


def cpu_port_rate(case, y, w_dose, w_grad, seconds=10.0, min_evals=3, threads=0):
    """Evaluations/s of the C oracle port (float64, OpenMP over rows) on the host cores, plus the
    scipy.sparse single-thread path the reference itself runs, on the full case."""
    from oracle.c_oracle import COracle
    co = COracle(case)
    nthreads = threads or (os.cpu_count() or 1)
    co.eval(y, w_dose, w_grad, threads=nthreads)
    n, t0 = 0, time.perf_counter()
    while True:
        co.eval(y, w_dose, w_grad, threads=nthreads)
        n += 1
        el = time.perf_counter() - t0
        if (el >= seconds and n >= min_evals) or n >= 2000:
            break
    rate = n / el
    # scipy path: CSC scatter D.x, numpy penalty, CSR D^T.vg - 1 core, as the reference executes
    from oracle import vmat_oracle as O
    import scipy.sparse as sp
    Dc = sp.csc_matrix(case.D)
    Dt = sp.csr_matrix(case.D.T)
    x = np.repeat(y, np.diff(case.beam_ptr)) * w_dose
    m, t1 = 0, time.perf_counter()
    while m < 3:
        d = Dc @ x
        f, vg = O.penalty(d, case.thr, case.over, case.under)
        gb = Dt @ vg
        m += 1
    scipy_rate = m / (time.perf_counter() - t1)
    return rate, nthreads, n, scipy_rate


def run_reference(args):
    """Reference arm: the path on the host cores.  The reference is pure Python on scipy.sparse
    and cannot be installed or run as a package, so this times the oracle port of
    the same arithmetic (oracle/vmat_oracle.c, float64, all host threads) on the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case, y, w_dose, w_grad = workload(args.workload, args.scale)
    from oracle.c_oracle import COracle
    co = COracle(case)
    threads = os.cpu_count() or 1
    for _ in range(max(1, min(args.warmup, 3))):
        co.eval(y, w_dose, w_grad, threads=threads)
    steps = min(args.steps, 50)
    t0 = time.perf_counter()
    for _ in range(steps):
        co.eval(y, w_dose, w_grad, threads=threads)
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {
        "impl": "reference", "metric": "dose+gradient evaluations/s", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {case.V} voxels x {case.B} beamlets, {case.nb} control points, "
                               f"nnz {case.nnz}", "scale": args.scale},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} full evaluations of the same case, oracle/vmat_oracle.c (OpenMP)"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=float, default=1.0, help="scale the voxel count (debugging only)")
    ap.add_argument("--store", default="f32")
    ap.add_argument("--acc", default="f32")
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--k2-tasks", type=int, default=0)
    ap.add_argument("--k2-threads", type=int, default=0)
    ap.add_argument("--k1-threads", type=int, default=0)
    ap.add_argument("--index-bytes", type=int, default=0, help="4 forces int32 column indices (default: uint16 when they fit)")
    ap.add_argument("--claim-lead", type=int, default=0, help="dose kernel: stages before a super-slice's end at which the next is claimed (tuning; 0 = default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from vmatwpencode_b200.plan import DosePlan
    from vmatwpencode_b200.dist import ShardedEvaluator, slab_of

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W = max(3, args.warmup)
    K = args.steps

    case, y, w_dose, w_grad = workload(args.workload, args.scale, device=f"cuda:{local}")
    if args.workload in DEVICE_WORKLOADS:
        from vmatwpencode_b200.dist import slab_of_device
        D, thr, over, under = slab_of_device(case, rank, world)
        torch.cuda.empty_cache()
    else:
        rows, D, thr, over, under = slab_of(case, rank, world)
    plan = DosePlan(D, case.beam_ptr, thr, over, under, device=local, store=args.store, acc=args.acc,
                    tile_voxels=args.tile, k2_tasks=args.k2_tasks, k2_threads=args.k2_threads,
                    k1_threads=args.k1_threads, index_bytes=args.index_bytes, claim_lead=args.claim_lead)
    if args.workload in DEVICE_WORKLOADS:      # the CSR copies are no longer needed: the plan holds its own layouts
        case.csr = None
        del D
        torch.cuda.empty_cache()
    plan.set_weights(w_dose, w_grad)
    info = plan.info()
    ev = ShardedEvaluator(plan)
    # a side stream: the library treats stream 0 as "use the plan's own stream", and the CUDA
    # events below must sit on the stream the kernels are launched on
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    stream = side.cuda_stream
    assert stream != 0
    plan.set_intensities(y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident throughput (`value`): y already in HBM, results left in HBM ----------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        ev.eval_async(stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        ev.eval_async(stream)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # second pass of the same K steps with a CUDA event after every kernel: the per-kernel
    # durations behind `roofline` (events between kernels serialise the launches, so this pass
    # is a little slower than the one `value` comes from)
    kt = None
    if world == 1:
        plan.set_timing(True)
        for _ in range(K):
            ev.eval_async(stream)
        barrier()
        kt = plan.get_timing()
        plan.set_timing(False)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    f_res, g_res = plan.fetch_result()

    # ---- end to end (`e2e`): host y in, host (f, g) out through the C ABI each step --------
    for _ in range(W):
        ev.eval(y) if world > 1 else plan.eval(y)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        yk = y if (k & 1) == 0 else y * 1.0000001
        f_e, g_e = ev.eval(yk) if world > 1 else plan.eval(yk)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_src = measured_peaks()
        value = K / (ms * 1e-3)
        sv = {"f32": 4, "bf16": 2, "f64": 8}[args.store]
        nnz_local = info["nnz"]
        line = {
            "metric": "dose+gradient evaluations/s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.acc, "data": "synthetic",
            "config": {"workload": f"{args.workload}: {case.V} voxels x {case.B} beamlets, {case.nb} control points, "
                                   f"nnz {case.nnz} ({100.0 * case.nnz / (case.V * case.B):.2f} % dense), geometric sparsity",
                       "nnz_this_rank": nnz_local,
                       "store_dtype": args.store, "acc_dtype": args.acc, "index_bytes": info["index_bytes"], "sharding": f"voxel slabs x{world}",
                       "l2": "both layouts of D (2 x ~%d MB) exceed the 126 MB L2; no flush between steps"
                             % (nnz_local * (sv + info["index_bytes"]) // 1_000_000),
                       "scale": args.scale},
            "e2e": {"value": K / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": case.nb * 8,
                    "d2h_bytes_per_step": (case.nb + 1) * 8, "ms_per_step": 1e3 * e2e_s / K,
                    "how": "vmat_eval(y, &f, g) per step with host buffers: y (nb f64) goes to the device inside the "
                           "dose kernel's launch parameters (a 1.4 KB H2D copy when nb > 384), f and g (nb+1 f64) come "
                           "back with one D2H copy into pinned memory and a stream synchronisation"},
            "gpu_launches": (info["launches_per_eval"] + (1 if world > 1 else 0)) * K,
            "clocks": clocks,
            "check": {"f": f_res, "g_norm": float(np.linalg.norm(g_res))},
        }
        algo = info["algorithmic_bytes_per_eval"]
        if kt is not None and kt["n"] > 0:
            # dominant kernel: whichever of K1 (dose+penalty) / K2 (transposed) is slower
            si = info["index_bytes"]
            k1_bytes = nnz_local * (sv + si) + 4 * (plan.V + 1) + 12 * plan.V + 8 * plan.V + 4 * plan.B
            k2_bytes = nnz_local * (sv + si) + 4 * plan.V + 4 * (plan.B + 1) + 4 * plan.B
            name, kms, kb = ("k2_beamlet_grad", kt["k2_ms"], k2_bytes) if kt["k2_ms"] >= kt["k1_ms"] else \
                            ("k1_dose_penalty", kt["k1_ms"], k1_bytes)
            ach = kb / (kms * 1e-3) / 1e9
            traffic = None      # dram__bytes_read+write of that kernel from the committed ncu --set full capture
            tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
            if os.path.isfile(tpath) and args.workload == "cfg2" and args.scale == 1.0 and \
                    (args.store, args.acc, info["index_bytes"]) == ("f32", "f32", json.load(open(tpath)).get("index_bytes", 4)):
                traffic = json.load(open(tpath))["dram_bytes_per_launch"].get(name)
            line["roofline"] = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s",
                                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                                "algorithmic_bytes_per_launch": kb, "launch_ms": kms}
            line["kernels_ms"] = {k: kt[k] for k in ("k1_ms", "k2_ms", "k3_ms", "eval_ms")}
            ach_eval = algo / (ms / K * 1e-3) / 1e9
            line["roofline_eval"] = {"achieved": ach_eval, "peak": peak, "unit": "GB/s", "frac": ach_eval / peak,
                                     "frac_of_8TBps": ach_eval / 8000.0, "algorithmic_bytes_per_eval": algo,
                                     "layout_bytes_per_eval": info["layout_bytes_per_eval"]}
        if world == 1 and not args.no_cpu and args.workload not in DEVICE_WORKLOADS:
            rate, cores, n, scipy_rate = cpu_port_rate(case, y, w_dose, w_grad)
            line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} full evaluations of the same case (oracle/vmat_oracle.c, f64, OpenMP)",
                                    "scipy_1core_value": scipy_rate}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    plan.close()


if __name__ == "__main__":
    main()
