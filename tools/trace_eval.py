#!/usr/bin/env python
"""Per-CTA timeline of one evaluation (diagnostics): when each CTA of the dose kernel and of the
transposed kernel entered, had its inputs, saw its first stage, and finished.  Prints a summary."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmatwpencode_b200.plan import DosePlan
from vmatwpencode_b200.synth import make_config


def summary(name, t):
    us = lambda a: [round(float(v) / 1e3, 2) for v in a]
    q = lambda col: us(np.percentile(t[:, col][t[:, col] >= 0], [0, 50, 100])) if (t[:, col] >= 0).any() else None
    return {name: dict(ctas=int(t.shape[0]), entry_us=q(0), inputs_ready_us=q(1), first_stage_us=q(2),
                       consumers_done_us=q(3), producer_done_us=q(4))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--index-bytes", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dump", default="", help="write the raw stamps of every repetition to this .npz")
    args = ap.parse_args()
    case = make_config(args.workload)
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, index_bytes=args.index_bytes)
    rng = np.random.default_rng(1)
    plan.set_weights(rng.random(case.B), rng.random(case.B))
    y = rng.random(case.nb)
    for _ in range(20):
        plan.eval(y)
    raw = {}
    for rep in range(args.reps):
        k1, k2 = plan.trace_eval()
        raw[f"k1_{rep}"], raw[f"k2_{rep}"] = k1, k2
        out = dict(workload=args.workload, unit="us since the first stamp; [min, median, max] over CTAs")
        out.update(summary("k1", k1))
        out.update(summary("k2", k2))
        print(json.dumps(out))
    if args.dump:
        np.savez(args.dump, **raw)
    plan.close()


if __name__ == "__main__":
    main()
