#!/usr/bin/env python
"""BASELINE.json configs[2] as written: the greedy loop with k-medoid VOXEL subsampling on the ~1M-voxel
phantom, sharded over the GPUs of one node (run under torchrun; works with 1 GPU too).

  1. every rank generates the 1M-voxel x 14 580-beamlet case on its GPU (device generator, config 3);
  2. k voxels are chosen by greedy k-medoid insertion over the voxels' grid coordinates
     (kmedoids.py:32-68 generalised to points in space): every insertion step prices all candidate voxels
     against ALL 1M voxels; the distance pass (K6) is sharded over the ranks' shares of the voxels and the
     candidate costs are all-reduced;
  3. the case is restricted to the sampled voxels (synth.subsample_device_case) and bound voxel-sharded
     (VMATlibrary.bind(shard=True)): every rank holds a slab of the subsample, evaluations exchange the
     beamlet gradient inside the reduction kernel;
  4. colGen runs in lock step on all ranks; rank 0 compares the aperture sequence with the CPU oracle on the
     same subsample.
Prints one JSON line (rank 0): wall time per phase, K6 distance pairs/s per GPU, the sequence check.
(Lives under tests/ because rank 0 runs the oracle as the checker; pytest does not collect it.)"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=2048, help="voxels to sample")
    ap.add_argument("--cand-stride", type=int, default=31, help="every n-th voxel is a candidate medoid (1 = all)")
    ap.add_argument("--iters", type=int, default=12, help="greedy iterations")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--single-steps", type=int, default=3, help="insertion steps timed unsharded on rank 0 for the K6 scaling figure")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = dist.group.WORLD if world > 1 else None
    from vmatwpencode_b200 import VMATlibrary as vl, kmedoids as km
    from vmatwpencode_b200.synth import make_device_config, voxel_coordinates, subsample_device_case

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    t0 = time.perf_counter()
    case = make_device_config("cfg3", scale=args.scale, device=f"cuda:{local}")
    sync()
    t_gen = time.perf_counter() - t0
    xyz = voxel_coordinates(case.V, case.M)
    cand = np.arange(0, case.V, max(1, args.cand_stride), dtype=np.int64)
    start = int(cand[len(cand) // 2])

    # K6 on one GPU (rank 0, a few insertion steps) for the scaling figure
    single = None
    if rank == 0 and args.single_steps > 0:
        from vmatwpencode_b200.plan import KMedoidSession
        cur = np.sqrt(((xyz - xyz[start]) ** 2).sum(axis=1))
        ses = KMedoidSession(xyz, cur, np.ascontiguousarray(xyz[cand]), local)     # the whole point set on one GPU
        ses.costs()                                                                 # warm-up (module load, clocks)
        t1 = time.perf_counter()
        for _ in range(args.single_steps):
            ses.costs()
        single = (time.perf_counter() - t1) / args.single_steps
        ses.close()
    sync()

    t1 = time.perf_counter()
    rows = km.sample_voxels(xyz, args.k, start=start, device=local, group=group, candidates=cand)
    sync()
    t_sample = time.perf_counter() - t1
    steps = len(rows) - 1
    pairs = float(len(cand)) * case.V * steps

    t1 = time.perf_counter()
    sub = subsample_device_case(case, rows)
    case.csr = None
    torch.cuda.empty_cache()
    data = vl.bind(sub, device=local, store="f32", acc="f64", group=group, shard=world > 1)
    sync()
    t_bind = time.perf_counter() - t1

    t1 = time.perf_counter()
    vl.colGen(0.01, False, 16, max_iter=args.iters)
    sync()
    t_loop = time.perf_counter() - t1
    hist = [(h["idx"], h["l"], h["r"]) for h in data.history]
    rmp = sum(h["nfev"] for h in data.history)
    same_on_all = True
    if world > 1:
        allh = [None] * world
        dist.all_gather_object(allh, hist)
        same_on_all = all(h == allh[0] for h in allh)

    if rank == 0:
        from oracle import vmat_oracle as O
        t1 = time.perf_counter()
        want = O.col_gen(O.OracleState(sub), C=0.01, kappasize=16, max_iter=args.iters, clean=True)
        t_oracle = time.perf_counter() - t1
        ok = hist == [(h["idx"], h["l"], h["r"]) for h in want]
        line = dict(workload=f"config 3: {case.V} voxels x {case.B} beamlets, {case.nb} control points, nnz {case.nnz}",
                    n_gpus=world, sampled_voxels=len(rows), candidates=len(cand), candidate_stride=args.cand_stride,
                    seconds=dict(generate=round(t_gen, 3), kmedoid_sampling=round(t_sample, 3), subsample_and_bind=round(t_bind, 3),
                                 greedy_loop=round(t_loop, 3), oracle_greedy_loop_cpu=round(t_oracle, 3)),
                    kmedoid=dict(insertion_steps=steps, distance_pairs=pairs, pairs_per_s_total=pairs / t_sample,
                                 pairs_per_s_per_gpu=pairs / t_sample / world, seconds_per_step=t_sample / max(1, steps),
                                 single_gpu_seconds_per_step=single,
                                 speedup_vs_one_gpu=(single / (t_sample / max(1, steps))) if single else None),
                    greedy=dict(iterations=len(hist), evaluations_in_rmps=rmp, sequence=[h[0] for h in hist],
                                sub_nnz=int(sub.D.nnz), identical_on_all_ranks=same_on_all,
                                aperture_sequence_equals_oracle=bool(ok)))
        print(json.dumps(line))
    data.plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
