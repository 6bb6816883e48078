#!/usr/bin/env python
"""Tuning sweep: build the bench workload once, then time plan variants with the library's
per-kernel CUDA events.  Prints one line per variant (ms per kernel, GB/s per kernel)."""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from bench import workload
from vmatwpencode_b200.plan import DosePlan


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--variants", default="")
    args = ap.parse_args()
    case, y, w_dose, w_grad = workload(args.workload, args.scale)
    Dt = case.D.tocsc()
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    variants = json.loads(args.variants) if args.variants else [
        dict(), dict(k2_stages=12), dict(k2_stages=16), dict(tile_voxels=4096), dict(tile_voxels=4096, k2_stages=12),
        dict(tile_voxels=4096, k2_stages=16), dict(tile_voxels=2048, k2_stages=16), dict(k1_stages=8), dict(k1_stages=16),
        dict(index_bytes=4), dict(store="bf16"), dict(acc="f64"),
    ]
    for v in variants:
        v = dict(v)
        store = v.pop("store", "f32"); acc = v.pop("acc", "f32")
        plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, store=store, acc=acc, Dt=Dt, **v)
        plan.set_weights(w_dose, w_grad)
        plan.set_intensities(y)
        info = plan.info()
        for _ in range(20):
            plan.eval_async(side.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            plan.eval_async(side.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        pipelined_ms = e0.elapsed_time(e1) / args.steps        # back to back, kernels overlapping by PDL
        import time
        t0 = time.perf_counter()
        for _ in range(args.steps):
            plan.eval(y)
        e2e_ms = 1e3 * (time.perf_counter() - t0) / args.steps
        k1t, _ = plan.trace_eval()
        done = k1t[:, 3] / 1e3
        plan.set_timing(True)
        e0.record()
        for _ in range(args.steps):
            plan.eval_async(side.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        t = plan.get_timing()
        ms = e0.elapsed_time(e1) / args.steps
        sv = {"f32": 4, "bf16": 2, "f64": 8}[store]
        kb = info["nnz"] * (sv + info["index_bytes"]) / 1e9
        print(json.dumps(dict(variant=dict(v, store=store, acc=acc), pipelined_ms=round(pipelined_ms, 5), e2e_ms=round(e2e_ms, 5),
                              k1_cta_done_us=[round(float(q), 1) for q in np.percentile(done, [0, 50, 100])], step_ms=round(ms, 5),
                              k1_ms=round(t["k1_ms"], 5), k2_ms=round(t["k2_ms"], 5), k3_ms=round(t["k3_ms"], 5),
                              k1_GBps=round(kb / t["k1_ms"] * 1e3, 1), k2_GBps=round(kb / t["k2_ms"] * 1e3, 1),
                              eval_GBps=round(info["algorithmic_bytes_per_eval"] / ms / 1e6, 1),
                              layout_MB=round(info["layout_bytes_per_eval"] / 1e6, 1))), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
