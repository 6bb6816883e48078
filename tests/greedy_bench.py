#!/usr/bin/env python
"""End-to-end greedy column generation on BASELINE.json configs[0] (20k voxels x 2 016 beamlets,
36 control points): wall time of `iters` outer iterations (evaluate, price all candidates, add
the best aperture, re-optimise with scipy L-BFGS-B) on the GPU path - with the sparse kernels
only and with solveRMC's dense column cache - against the CPU oracle running the same loop
(reference arithmetic restated; 'literal' is the reference's own operation order).  The aperture
sequences must be identical.  Prints one JSON line.  (Lives under tests/ because it runs the
oracle as the CPU side of the comparison; pytest does not collect it.)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import vmat_oracle as O
from vmatwpencode_b200 import VMATlibrary as vl
from vmatwpencode_b200.synth import make_config


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg1")
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--C", type=float, default=0.01)
    ap.add_argument("--kappa", type=int, default=16)
    ap.add_argument("--profile", action="store_true", help="cProfile the GPU loop (sparse only) and print the top of it to stderr")
    ap.add_argument("--literal", action="store_true", help="also time the reference's literal operation order (slow)")
    args = ap.parse_args()
    case = make_config(args.workload)
    out = dict(workload=f"{args.workload}: {case.V} voxels x {case.B} beamlets, {case.nb} control points",
               iters=args.iters, C=args.C, kappasize=args.kappa, host_cores=os.cpu_count())
    seqs = {}
    for name, flag, lean in (("gpu_column_cache", True, False), ("gpu_sparse_minimize", False, False), ("gpu_sparse_only", False, True)):
        vl.use_column_cache = flag
        vl.fast_rmp = lean             # gpu_sparse_only: the default path (lean loop over scipy's compiled L-BFGS-B routine)
        data = vl.bind(case, store="f32", acc="f64")
        t0 = time.perf_counter()
        vl.colGen(args.C, False, args.kappa, max_iter=args.iters)
        out[name + "_s"] = round(time.perf_counter() - t0, 4)
        out[name + "_evals"] = sum(h["nfev"] + 2 for h in data.history)
        out[name + "_rmp_s"] = round(sum(h.get("rmp_s", 0.0) for h in data.history), 4)
        seqs[name] = [(h["idx"], tuple(h["l"]), tuple(h["r"])) for h in data.history]
        data.plan.close()
    vl.use_column_cache = False
    vl.fast_rmp = True
    if args.profile:
        import cProfile
        import pstats
        data = vl.bind(case, store="f32", acc="f64")
        pr = cProfile.Profile()
        pr.enable()
        vl.colGen(args.C, False, args.kappa, max_iter=args.iters)
        pr.disable()
        data.plan.close()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    t0 = time.perf_counter()
    hist = O.col_gen(O.OracleState(case), C=args.C, kappasize=args.kappa, max_iter=args.iters, clean=True)
    out["cpu_oracle_clean_s"] = round(time.perf_counter() - t0, 4)
    seqs["cpu"] = [(h["idx"], tuple(h["l"]), tuple(h["r"])) for h in hist]
    if args.literal:
        t0 = time.perf_counter()
        hist2 = O.col_gen(O.OracleState(case), C=args.C, kappasize=args.kappa, max_iter=args.iters)
        out["cpu_reference_order_s"] = round(time.perf_counter() - t0, 4)
        seqs["cpu_literal"] = [(h["idx"], tuple(h["l"]), tuple(h["r"])) for h in hist2]
    out["identical_aperture_sequences"] = all(s == seqs["cpu"] for s in seqs.values())
    out["sequence"] = [s[0] for s in seqs["cpu"]]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
