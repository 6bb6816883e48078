#!/usr/bin/env python
"""bench.py - dose+gradient evaluations per second on BASELINE.json's configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2]

One step = one evaluation of the hot path: x = y*w_dose, d = D x, penalty (objective f and
voxel gradient), gb = D^T vg, aperture gradient g  (calcObjGrad, greedyVMATsampling.py:578-582
plus the beamlet gradient PPsubroutine :646 needs).  Workload at N=1: BASELINE.json
configs[1], the synthetic 300k-voxel x 10 080-beamlet x 180-control-point case at 1 % density.
For N>1 the same case is voxel-sharded (strong scaling) with one exchange per evaluation, fused
into the reduction kernel over NVLink peer memory.  After the timed legs the results are compared
with the CPU oracle (`check`), and the largest case (BASELINE.json configs[3], 4M voxels,
device-generated, voxel-sharded) is measured and checked in the same process (`largest_case`).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

L2_BYTES = 126_000_000


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled in-process through NVML on a thread while
    the timed regions run (a 20-step run is over before `nvidia-smi -lms` prints its first line)."""
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, index=0):
        self.index, self.sm, self.bits, self.label = index, {}, {}, "headline"
        self.stop_flag = threading.Event()
        self.loaded = threading.Event()          # set while a timed region is running
        self.nv = self.h = self.t = None
        self.smax = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.nv = None

    def _run(self):
        nv, h = self.nv, self.h
        while not self.stop_flag.is_set():
            if self.loaded.is_set():
                try:
                    lab = self.label
                    self.sm.setdefault(lab, []).append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                    self.bits[lab] = self.bits.get(lab, 0) | int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    pass
                time.sleep(0.0002)
            else:
                time.sleep(0.0005)

    def stop(self):
        if self.nv is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)

    def summary(self, label):
        """Clocks seen while the timed regions of workload `label` ran."""
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        sm, bits = self.sm.get(label, []), self.bits.get(label, 0)
        reasons = [n for n, b in self.REASONS if bits & b]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                "samples": len(sm), "how": "NVML, sampled in-process only while this workload's timed regions run"}


DEVICE_WORKLOADS = ("cfg3", "cfg4")      # too large to build through scipy on the host


def describe(name, V, B, nb, nnz):
    """config.workload - the same string in both arms."""
    return f"{name}: {V} voxels x {B} beamlets, {nb} control points, nnz {nnz}, 1 % dense geometric sparsity"


def bench_weights(nb, M, N, beam_ptr, B):
    """All control points selected with a centred open aperture; intensities in (0.5, 1.5)."""
    w_dose = np.zeros(B)
    w_grad = np.zeros(B)
    for i in range(nb):
        lo = int(beam_ptr[i])
        for m in range(M):
            w_dose[lo + m * N + 1: lo + (m + 1) * N - 1] = 1.0
            w_grad[lo + m * N + 1: lo + (m + 1) * N - 1] = 1.0
    y = 0.5 + (np.arange(nb) * 0.6180339887) % 1.0
    return y, w_dose, w_grad


def workload(name, scale, device=None):
    from vmatwpencode_b200.synth import make_config, make_device_config
    if name in DEVICE_WORKLOADS and device is not None:
        case = make_device_config(name, scale=scale, device=device)
    else:
        case = make_config(name, scale=scale)
    y, w_dose, w_grad = bench_weights(case.nb, case.M, case.N, case.beam_ptr, case.B)
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
    return rate, nthreads, n, scipy_rate, co


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
    warm = max(1, min(args.warmup, 3))            # a CPU evaluation is ~5 ms: a few are enough to warm caches
    for _ in range(warm):
        co.eval(y, w_dose, w_grad, threads=threads)
    steps = max(1, min(args.steps, 50))
    t0 = time.perf_counter()
    for _ in range(steps):
        co.eval(y, w_dose, w_grad, threads=threads)
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {
        "impl": "reference", "metric": "dose+gradient evaluations/s", "value": val, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "steps_requested": args.steps,
        "warmup_requested": args.warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": describe(args.workload, case.V, case.B, case.nb, case.nnz), "scale": args.scale},
        "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": f"{steps} full evaluations of the same case, oracle/vmat_oracle.c (OpenMP)"},
        "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def relerr(a, b):
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    den = float(np.abs(b).max()) if b.size else 0.0
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0)) if a.size else 0.0


def sampled_device_check(plan, csr, thr, over, under, y, w_dose, w_grad, beam_ptr, f_dev, g_dev, rank, world, dist, torch,
                         nrows=10000, ncols=1000):
    """Check of a device-generated (possibly sharded) case without a host copy of D: sampled voxel rows of
    the dose and sampled beamlets of the beamlet gradient are recomputed in float64 with torch index
    arithmetic from the rank's CSR tensors (not the plan's layouts, not the plan's kernels), the
    objective from the device dose, the aperture gradient from the device beamlet gradient."""
    dev = csr["vval"].device
    V = int(csr["vrowptr"].numel()) - 1
    B = int(beam_ptr[-1])
    x = torch.as_tensor(np.repeat(y, np.diff(beam_ptr)) * w_dose, device=dev)
    d_dev = torch.as_tensor(plan.dose(), device=dev)
    vg_dev = torch.as_tensor(plan.voxelgrad(), device=dev)
    gb_dev = plan.beamlet_grad()
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)

    def rows_dot(rowptr, col, val, rows, vec):
        lo, hi = rowptr[rows], rowptr[rows + 1]
        n = hi - lo
        seg = torch.repeat_interleave(torch.arange(len(rows), device=dev), n)
        pos = torch.arange(int(n.sum().item()), device=dev) - torch.repeat_interleave(torch.cumsum(n, 0) - n, n) + \
            torch.repeat_interleave(lo, n)
        out = torch.zeros(len(rows), dtype=torch.float64, device=dev)
        out.index_add_(0, seg, val[pos].double() * vec[col[pos].long()])
        return out

    rows = torch.randperm(V, generator=gen)[:min(nrows, V)].to(dev)
    d_ref = rows_dot(csr["vrowptr"], csr["vcol"], csr["vval"], rows, x)
    t = torch.as_tensor(thr, device=dev)[rows]
    ov = torch.as_tensor(over, device=dev)[rows]
    un = torch.as_tensor(under, device=dev)[rows]
    o = torch.clamp(d_ref - t, min=0)
    u = torch.clamp(t - d_ref, min=0)
    vg_ref = 2 * (2 * o * ov - 2 * u * un)
    err_d = float((d_dev[rows] - d_ref).abs().max() / d_dev.abs().max().clamp(min=1e-300))
    err_vg = float((vg_dev[rows] - vg_ref).abs().max() / vg_dev.abs().max().clamp(min=1e-300))
    # objective from the whole device dose (this rank's slab), summed over ranks
    tt, oo, uu = (torch.as_tensor(a, device=dev) for a in (thr, over, under))
    o = torch.clamp(d_dev - tt, min=0)
    u = torch.clamp(tt - d_dev, min=0)
    f_ref = (o * o * oo + u * u * uu).sum().reshape(1)
    cols = torch.randperm(B, generator=torch.Generator(device="cpu").manual_seed(99))[:min(ncols, B)].to(dev)
    gb_ref = rows_dot(csr["browptr"], csr["bcol"], csr["bval"], cols, vg_dev)
    if world > 1:
        dist.all_reduce(f_ref)
        dist.all_reduce(gb_ref)
    gb_s = torch.as_tensor(gb_dev, device=dev)
    err_gb = float((gb_s[cols] - gb_ref).abs().max() / gb_s.abs().max().clamp(min=1e-300))
    g_ref = np.add.reduceat(gb_dev * w_grad, np.asarray(beam_ptr[:-1], dtype=np.int64))
    errs = torch.tensor([err_d, err_vg], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    return {"rel_err_f": abs(f_dev - float(f_ref.item())) / abs(float(f_ref.item())),
            "rel_err_g": relerr(g_dev, g_ref), "rel_err_gb": err_gb,
            "rel_err_d": float(errs[0].item()), "rel_err_vg": float(errs[1].item()),
            "against": f"float64 torch recomputation from the ranks' CSR tensors: {len(rows)} random voxel rows of d and vg "
                       f"per rank, {len(cols)} random beamlets of gb (summed over ranks), f from the device dose, "
                       "g from the device gb"}


def run_case(args, name, K, W, rank, local, world, dist, torch, sampler, opts, headline):
    """Build the plan of workload `name` on this rank's slab, time the resident and the host-buffer legs,
    time the kernels, check the results.  Returns the measurement dict (meaningful on rank 0)."""
    from vmatwpencode_b200.plan import DosePlan
    from vmatwpencode_b200.dist import ShardedEvaluator, slab_of, slab_of_device
    device_case = name in DEVICE_WORKLOADS
    sampler.label = "headline" if headline else name
    case, y, w_dose, w_grad = workload(name, args.scale, device=f"cuda:{local}")
    if device_case:
        D, thr, over, under = slab_of_device(case, rank, world)
        if world > 1:
            case.csr = None
        torch.cuda.empty_cache()
    else:
        rows, D, thr, over, under = slab_of(case, rank, world)
    plan = DosePlan(D, case.beam_ptr, thr, over, under, device=local, store=args.store, acc=args.acc, **opts)
    plan.set_weights(w_dose, w_grad)
    info = plan.info()
    ev = ShardedEvaluator(plan)
    stream = torch.cuda.current_stream().cuda_stream
    assert stream != 0
    plan.set_intensities(y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident throughput (`value`): y already in HBM, results left in HBM ----------
    for _ in range(W):
        ev.eval_async(stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.loaded.set()
    e0.record()
    for _ in range(K):
        ev.eval_async(stream)
    e1.record()
    barrier()
    sampler.loaded.clear()
    ms = maxr(e0.elapsed_time(e1))
    # second pass of the same K steps with a CUDA event after every kernel: the per-kernel
    # durations behind `roofline` (events between kernels serialise the launches, so this pass
    # is a little slower than the one `value` comes from)
    plan.set_timing(True)
    sampler.loaded.set()
    for _ in range(K):
        ev.eval_async(stream)
    barrier()
    sampler.loaded.clear()
    kt = plan.get_timing()
    plan.set_timing(False)
    f_res, g_res = plan.fetch_result()

    # ---- end to end (`e2e`): host y in, host (f, g) out through the C ABI each step --------
    We = max(W, 50)              # the host side of this leg (ctypes call, pinned buffers, the driver's launch path) warms up too
    for _ in range(We):
        plan.eval(y)
    barrier()
    sampler.loaded.set()
    t0 = time.perf_counter()
    for k in range(K):
        yk = y if (k & 1) == 0 else y * 1.0000001
        plan.eval(yk)
    barrier()
    e2e_s = maxr(time.perf_counter() - t0)
    sampler.loaded.clear()
    if headline:
        # sustained leg for the clock sample only (not reported as a rate): ~0.2 s of back-to-back evaluations
        n_sus = max(K, int(0.2 / max(ms / K * 1e-3, 1e-6)))
        sampler.loaded.set()
        for _ in range(min(n_sus, 20000)):
            ev.eval_async(stream)
        barrier()
        sampler.loaded.clear()
        plan.set_intensities(y)
        ev.eval_async(stream)
        barrier()
        f_res, g_res = plan.fetch_result()

    trace_us = None
    if args.trace:          # per-CTA time stamps of one evaluation (every rank runs it: it includes the exchange)
        barrier()
        k1t, k2t = plan.trace_eval()
        q = lambda col: [round(float(v) / 1e3, 2) for v in np.percentile(col[col >= 0], [0, 50, 100])] if (col >= 0).any() else None
        if info["launches_per_eval"] == 1:
            names = ["entry", "x_staged", "first_stage", "first_transposed_item", "queue_empty", "barrier_passed", "end"]
            trace_us = {n: q(plan.trace_raw[:, i]) for i, n in enumerate(names)}
        else:
            trace_us = dict(k1_done=q(k1t[:, 3]), k2_entry=q(k2t[:, 0]), k2_done=q(k2t[:, 3]))
            k3t = plan.trace_k3
            if k3t is not None:      # reduction kernel: start, own partials reduced (and pushed), every rank's values in, done
                trace_us.update(k3_start=q(k3t[:, 0]), k3_own_partials_pushed=q(k3t[:, 1]), k3_all_ranks_summed=q(k3t[:, 2]),
                                k3_done=q(k3t[:, 3]))
        barrier()

    # ---- check (outside every timed region) ---------------------------------------------------
    check = {"f": f_res, "g_norm": float(np.linalg.norm(g_res))}
    if device_case:
        check.update(sampled_device_check(plan, D, thr, over, under, y, w_dose, w_grad, case.beam_ptr, f_res, g_res,
                                          rank, world, dist, torch))
    elif rank == 0:
        from oracle.c_oracle import COracle
        co = COracle(case)
        f0, g0 = co.eval(y, w_dose, w_grad, threads=os.cpu_count() or 1)
        check.update({"rel_err_f": abs(f_res - f0) / abs(f0), "rel_err_g": relerr(g_res, g0),
                      "rel_err_gb": relerr(plan.beamlet_grad(), co.gb),
                      "against": "oracle/vmat_oracle.c (float64) on the full case, after the timed legs"})
        if world == 1:
            check["rel_err_d"] = relerr(plan.dose(), co.d)
            check["rel_err_vg"] = relerr(plan.voxelgrad(), co.vg)
    errs = [v for k, v in check.items() if k.startswith("rel_err")]
    if errs:
        check["tolerance"] = 1e-5 if args.acc == "f32" else 1e-10
        check["ok"] = bool(max(errs) <= check["tolerance"])

    out = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        sv = {"f32": 4, "bf16": 2, "f64": 8}[args.store]
        si = info["index_bytes"]
        nnz_local = info["nnz"]
        slab_mb = nnz_local * (sv + si) / 1e6
        fits = 2 * slab_mb * 1e6 <= L2_BYTES
        out = {
            "workload": describe(name, case.V, case.B, case.nb, case.nnz if not device_case else int(case.nnz)),
            "value": K / (ms * 1e-3), "ms_per_step": ms / K, "steps": K, "warmup": W,
            "e2e": {"value": K / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": case.nb * 8,
                    "d2h_bytes_per_step": (case.nb + (2 if world > 1 else 1)) * 8, "ms_per_step": 1e3 * e2e_s / K,
                    "steps": K, "warmup": We,
                    "how": "vmat_eval(y, &f, g) per step with host buffers on every rank: y (nb f64) goes to the device "
                           "inside the dose kernel's launch parameters (a 1.4 KB H2D copy when nb > 384), f and g "
                           "(nb+1 f64) come back with one D2H copy into pinned memory and one stream synchronisation"},
            "plan": {"nnz_this_rank": nnz_local, "store_dtype": args.store, "acc_dtype": args.acc, "index_bytes": si,
                     "kernels_per_eval": info["launches_per_eval"],
                     "sharding": f"voxel slabs x{world}", "exchange": "push all-reduce fused into the reduction kernel "
                     "(NVLink peer memory)" if world > 1 else "none",
                     "l2": ("both layouts of this rank's slab (2 x %.0f MB) FIT the 126 MB L2: steps after the first "
                            "stream from L2, not HBM; no flush between steps" % slab_mb) if fits else
                           ("both layouts of this rank's slab (2 x %.0f MB) exceed the 126 MB L2; no flush between steps"
                            % slab_mb)},
            "gpu_launches": info["launches_per_eval"] * K,
            "check": check,
        }
        # dominant kernel of this rank: whichever of K1 (dose+penalty) / K2 (transposed) is slower
        if kt["n"] > 0:
            k1_bytes = nnz_local * (sv + si) + 4 * (plan.V + 1) + 12 * plan.V + 8 * plan.V + 4 * plan.B
            k2_bytes = nnz_local * (sv + si) + 4 * plan.V + 4 * (plan.B + 1) + 4 * plan.B
            kname, kms, kb = ("k2_beamlet_grad", kt["k2_ms"], k2_bytes) if kt["k2_ms"] >= kt["k1_ms"] else \
                             ("k1_dose_penalty", kt["k1_ms"], k1_bytes)
            if info["launches_per_eval"] == 1:      # the whole evaluation is one kernel: its bytes are the evaluation's
                kname, kms, kb = "kf_eval", kt["eval_ms"], info["algorithmic_bytes_per_eval"]
            ach = kb / (kms * 1e-3) / 1e9
            traffic = None      # dram__bytes_read+write of that kernel from the committed ncu --set full capture
            tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if os.path.isfile(tpath) and world == 1 and args.scale == 1.0:
                tj = json.load(open(tpath))
                key = f"{name}/{args.store}/{args.acc}/{si}"
                traffic = tj.get("dram_bytes_per_launch", {}).get(key, {}).get(kname)
            out["roofline"] = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": peak, "unit": "GB/s",
                               "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                               "algorithmic_bytes_per_launch": kb, "launch_ms": kms,
                               "scope": "rank 0's slab" if world > 1 else "whole case",
                               "bytes_per_nonzero": sv + si,
                               "note": ("slab fits L2: not an HBM measurement" if fits else
                                        ("byte-delta indices: %d algorithmic bytes per nonzero (uint16 indices: %d) - fewer "
                                         "bytes in less time, at a lower fraction of the peak than the uint16 layout "
                                         "reaches (DESIGN.md section 3)" % (sv + 1, sv + 2)) if si == 1 else None)}
            out["kernels_ms"] = {k: kt[k] for k in ("k1_ms", "k2_ms", "k3_ms", "eval_ms")}
            algo = info["algorithmic_bytes_per_eval"]
            ach_eval = algo / (ms / K * 1e-3) / 1e9
            out["roofline_eval"] = {"achieved": ach_eval, "peak": peak, "unit": "GB/s", "frac": ach_eval / peak,
                                    "frac_of_8TBps": ach_eval / 8000.0, "frac_of_read_peak_7090": ach_eval / 7090.0,
                                    "algorithmic_bytes_per_eval": algo,
                                    "layout_bytes_per_eval": info["layout_bytes_per_eval"],
                                    "scope": "rank 0's slab per evaluation" if world > 1 else "whole case"}
        if trace_us is not None:
            out["trace_us"] = trace_us
        out["_case"] = (case, y, w_dose, w_grad)
    barrier()
    plan.close()
    del plan, D
    if device_case:
        case.csr = None
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--index-bytes", type=int, default=0, help="index form: 1 byte deltas, 2 uint16 when it fits, 4 int32 (0 = the library's default: byte deltas for fp32/fp32 plans of 5 M nonzeros and more, else uint16)")
    ap.add_argument("--claim-lead", type=int, default=0, help="dose kernel: stages before a super-slice's end at which the next is claimed (tuning; 0 = default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fused", type=int, default=0, help="0 / 1 = the three kernels K1/K2/K3 (default), 2 = the single fused evaluation kernel")
    ap.add_argument("--trace", action="store_true", help="add per-CTA time stamps of one evaluation (diagnostics)")
    ap.add_argument("--largest", default="cfg4", help="workload of the `largest_case` block ('none' skips it)")
    ap.add_argument("--largest-steps", type=int, default=200)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

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
    opts = dict(tile_voxels=args.tile, k2_tasks=args.k2_tasks, k2_threads=args.k2_threads, k1_threads=args.k1_threads,
                index_bytes=args.index_bytes, claim_lead=args.claim_lead, fused=args.fused)
    # a side stream: the library treats stream 0 as "use the plan's own stream", and the CUDA
    # events must sit on the stream the kernels are launched on
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    head = run_case(args, args.workload, K, W, rank, local, world, dist, torch, sampler, opts, headline=True)
    largest = None
    if args.largest != "none" and args.largest != args.workload and args.scale == 1.0:
        largest = run_case(args, args.largest, max(1, args.largest_steps), W, rank, local, world, dist, torch, sampler,
                           opts, headline=False)
    sampler.stop()
    clocks = sampler.summary("headline") if rank == 0 else None

    if rank == 0:
        case, y, w_dose, w_grad = head.pop("_case")
        line = {
            "metric": "dose+gradient evaluations/s", "value": head["value"], "unit": "evals/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.acc, "data": "synthetic",
            "config": {"workload": head["workload"], "scale": args.scale},
            "plan": head["plan"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "clocks": clocks, "check": head["check"],
        }
        for k in ("roofline", "kernels_ms", "roofline_eval", "trace_us"):
            if k in head:
                line[k] = head[k]
        if largest is not None:
            largest.pop("_case")
            largest["scaling"] = "strong"
            largest["clocks"] = sampler.summary(args.largest)
            line["largest_case"] = largest
        if world == 1 and not args.no_cpu and args.workload not in DEVICE_WORKLOADS:
            rate, cores, n, scipy_rate, _ = cpu_port_rate(case, y, w_dose, w_grad)
            line["cpu_baseline"] = {"value": rate, "unit": "evals/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} full evaluations of the same case (oracle/vmat_oracle.c, f64, OpenMP)",
                                    "scipy_1core_value": scipy_rate}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
