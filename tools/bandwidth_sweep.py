#!/usr/bin/env python
"""BASELINE.json configs[4]: D.x / D^T.r bandwidth sweep over density 0.1-5 % and 100k-8M voxels
(B = 10 080 beamlets, uniform sparsity) against scipy.sparse on the host cores.

GPU: the two sparse kernels timed with CUDA events (library timing mode), algorithmic GB/s =
nnz*(4+4) bytes / kernel time.  CPU: scipy CSR gather D.x and CSR D^T.r on a bounded voxel sample
(float64 + int32, 12 bytes per nonzero, one core - scipy's mat-vec is single threaded).
Prints one JSON line per point."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import scipy.sparse as sp
import torch

from vmatwpencode_b200.plan import DosePlan
from vmatwpencode_b200.synth import make_device_case


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", default="100000,300000,1000000,4000000,8000000")
    ap.add_argument("--density", default="0.001,0.005,0.01,0.02,0.05")
    ap.add_argument("--max-nnz", type=float, default=1.3e9)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cpu-voxels", type=int, default=100000)
    args = ap.parse_args()
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    for V in [int(v) for v in args.voxels.split(",")]:
        for dens in [float(d) for d in args.density.split(",")]:
            nb, M, N = 180, 7, 8
            B = nb * M * N
            if V * B * dens > args.max_nnz:
                print(json.dumps(dict(V=V, density=dens, skipped="nnz above --max-nnz")), flush=True)
                continue
            case = make_device_case(V=V, nb=nb, M=M, N=N, density=dens, seed=1005, model="uniform")
            # CPU sample first (needs the CSR): the first cpu_voxels rows
            nv = min(V, args.cpu_voxels)
            p1 = int(case.csr["vrowptr"][nv].item())
            Dh = sp.csr_matrix((case.csr["vval"][:p1].cpu().numpy().astype(np.float64),
                                case.csr["vcol"][:p1].cpu().numpy(), case.csr["vrowptr"][:nv + 1].cpu().numpy()),
                               shape=(nv, B))
            Dth = sp.csr_matrix(Dh.T)
            x = np.random.default_rng(0).random(B); r = np.random.default_rng(1).random(nv)
            best_f = best_t = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); Dh @ x; best_f = min(best_f, time.perf_counter() - t0)
                t0 = time.perf_counter(); Dth @ r; best_t = min(best_t, time.perf_counter() - t0)
            plan = DosePlan(case.csr, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc="f32")
            nnz = case.nnz
            case.csr = None
            torch.cuda.empty_cache()
            w = np.ones(B)
            plan.set_weights(w, w)
            plan.set_intensities(np.ones(nb))
            for _ in range(5):
                plan.eval_async(side.cuda_stream)
            plan.set_timing(True)
            for _ in range(args.steps):
                plan.eval_async(side.cuda_stream)
            t = plan.get_timing()
            plan.set_timing(False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                plan.eval_async(side.cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            step_ms = e0.elapsed_time(e1) / args.steps
            info = plan.info()
            print(json.dumps(dict(V=V, density=dens, nnz=nnz, dose_kernel_GBps=round(nnz * 8 / t["k1_ms"] / 1e6, 1),
                                  transposed_kernel_GBps=round(nnz * 8 / t["k2_ms"] / 1e6, 1),
                                  eval_ms=round(step_ms, 4), evals_per_s=round(1e3 / step_ms, 1),
                                  eval_GBps=round(info["algorithmic_bytes_per_eval"] / step_ms / 1e6, 1),
                                  scipy_1core_Dx_GBps=round(Dh.nnz * 12 / best_f / 1e9, 2),
                                  scipy_1core_DTr_GBps=round(Dh.nnz * 12 / best_t / 1e9, 2),
                                  cpu_sample_voxels=nv, host_cores=os.cpu_count())), flush=True)
            plan.close()
            del plan
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
