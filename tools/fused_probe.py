#!/usr/bin/env python
"""Diagnostics for the fused evaluation kernel: per plan variant, the pipelined time per evaluation and the
per-CTA time stamps of one traced evaluation (us after the first CTA entered; [min, median, max] over CTAs):
x staged, first stage landed, first transposed item seen, queue empty, grid barrier passed, end."""
import argparse
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
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--variants", default="")
    ap.add_argument("--items", action="store_true", help="per-item stamps of three CTAs (fused kernel)")
    args = ap.parse_args()
    device_case = args.workload in ("cfg3", "cfg4")
    case, y, w_dose, w_grad = workload(args.workload, args.scale, device="cuda:0" if device_case else None)
    D = case.csr if device_case else case.D
    Dt = None if device_case else case.D.tocsc()
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    variants = json.loads(args.variants) if args.variants else [dict(fused=2), dict(fused=6), dict(fused=1)]
    for v in variants:
        v = dict(v)
        store = v.pop("store", "f32"); acc = v.pop("acc", "f32")
        kw = dict(Dt=Dt) if Dt is not None else {}
        plan = DosePlan(D, case.beam_ptr, case.thr, case.over, case.under, store=store, acc=acc, **kw, **v)
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
        ms = e0.elapsed_time(e1) / args.steps
        out = dict(variant=dict(v, store=store, acc=acc), us_per_eval=round(1e3 * ms, 2),
                   layout_MB=round(info["layout_bytes_per_eval"] / 1e6, 1), launches=info["launches_per_eval"],
                   GBps_layout=round(info["layout_bytes_per_eval"] / ms / 1e6, 1))
        k1, k2 = plan.trace_eval()
        us = lambda a: [round(float(q) / 1e3, 2) for q in np.percentile(a[a >= 0], [0, 50, 100])] if (a >= 0).any() else None
        if info["launches_per_eval"] == 1:
            raw = plan.trace_raw
            names = ["entry", "x_staged", "first_stage", "first_transposed_item", "queue_empty", "barrier_passed", "end"]
            out["trace_us"] = {n: us(raw[:, i]) for i, n in enumerate(names)}
            it = plan.trace_items
            if it is not None and args.items:
                t0 = int(it[it > 0].min())
                for cta in (0, 73, 147):
                    rows = []
                    for k in range(32):
                        r = it[cta, k]
                        if (r[:8] > 0).sum() == 0:
                            break
                        rows.append([round((int(v & ~1) - t0) / 1e3, 2) if v > 0 else -1 for v in r])
                    out[f"items_cta{cta}"] = rows
        else:
            out["trace_us"] = dict(k1_done=us(k1[:, 3]), k2_entry=us(k2[:, 0]), k2_done=us(k2[:, 3]))
            plan.set_timing(True)
            for _ in range(200):
                plan.eval_async(side.cuda_stream)
            torch.cuda.synchronize()
            out["kernels_us"] = {k: round(v * 1e3, 2) for k, v in plan.get_timing().items() if k.endswith("_ms")}
            plan.set_timing(False)
        print(json.dumps(out), flush=True)
        plan.close()


if __name__ == "__main__":
    main()
