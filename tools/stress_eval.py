#!/usr/bin/env python
"""Race hunt: many back-to-back evaluations of one plan, every result compared bit for bit with the first
(the kernels are deterministic, so any difference - or a CUDA fault - is a synchronisation bug).  One plan
variant per subprocess, because a fault poisons the CUDA context.

usage: stress_eval.py --workload cfg4 --evals 2000 --variants '[{}, {"index_bytes": 2}, {"consume_batch": 1}]'"""
import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(args, variant):
    import numpy as np
    import torch
    from bench import workload, DEVICE_WORKLOADS
    from vmatwpencode_b200.plan import DosePlan
    v = dict(variant)
    store, acc = v.pop("store", "f32"), v.pop("acc", "f32")
    case, y, w_dose, w_grad = workload(args.workload, args.scale, device="cuda:0")
    D = case.csr if args.workload in DEVICE_WORKLOADS else case.D
    plan = DosePlan(D, case.beam_ptr, case.thr, case.over, case.under, store=store, acc=acc, **v)
    plan.set_weights(w_dose, w_grad)
    plan.set_intensities(y)
    stream = torch.cuda.Stream()
    out = dict(variant=variant, index_bytes=plan.info()["index_bytes"], evals=0, mismatches=0)
    ref = None
    try:
        done = 0
        while done < args.evals:
            n = min(args.batch, args.evals - done)
            for _ in range(n):
                plan.eval_async(stream.cuda_stream)
            torch.cuda.synchronize()
            f, g = plan.fetch_result()
            gb = plan.beamlet_grad()
            cur = (f, g.tobytes(), gb.tobytes())
            if ref is None:
                ref = cur
            elif cur != ref:
                out["mismatches"] += 1
            done += n
            out["evals"] = done
        for k in range(args.host_evals):                      # the host-buffer path as well
            f, g = plan.eval(y)
            if (f, g.tobytes()) != ref[:2]:
                out["mismatches"] += 1
        out["host_evals"] = args.host_evals
        out["ok"] = out["mismatches"] == 0
    except Exception as e:                                    # noqa: BLE001 - report whatever the fault was
        out["ok"] = False
        out["error"] = str(e)[:200]
        try:
            out["guard"] = plan.guard_status()
        except Exception as e2:                               # noqa: BLE001
            out["guard"] = "unreadable: " + str(e2)[:100]
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--evals", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=100, help="evaluations between result comparisons")
    ap.add_argument("--host-evals", type=int, default=200)
    ap.add_argument("--variants", default="[{}]")
    ap.add_argument("--child", default="")
    args = ap.parse_args()
    if args.child:
        child(args, json.loads(args.child))
        return
    for v in json.loads(args.variants):
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--scale", str(args.scale),
               "--evals", str(args.evals), "--batch", str(args.batch), "--host-evals", str(args.host_evals),
               "--child", json.dumps(v)]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(lines[-1] if lines else json.dumps(dict(variant=v, ok=False, error=(r.stderr or "no output")[-300:])), flush=True)


if __name__ == "__main__":
    main()
