"""Voxel-slab sharding of one planning case across the GPUs of a node.

D is partitioned by voxel rows into contiguous slabs balanced by nonzeros; rank k owns the
slab's rows of D (in both layouts), its slice of the penalty tables, and computes a partial
beamlet gradient and objective.  One exchange per evaluation: a sum all-reduce of the
[gb | f] buffer (B+1 doubles) over NCCL; the aperture gradient is then formed redundantly on
every rank.  The reference has no multi-device path (its only parallelism is a
process pool over pricing candidates, greedyVMATsampling.py:836-841).
"""
from __future__ import annotations

import numpy as np


def partition_voxels(indptr, nparts: int):
    """Slab boundaries [v_0=0, ..., v_nparts=V] with ~equal nonzeros per slab, multiples of 32
    (so slabs keep whole slices) except the last.  Every slab gets at least 32 voxels: a case too small
    for that raises here, on every rank alike and before any collective, instead of leaving one rank
    with an empty plan while its peers wait in a barrier."""
    indptr = np.asarray(indptr, dtype=np.int64)
    V = len(indptr) - 1
    if V < 32 * nparts:
        raise ValueError(f"{V} voxels cannot be sharded over {nparts} ranks (at least 32 voxels per slab)")
    nnz = int(indptr[-1])
    bounds = [0]
    for k in range(1, nparts):
        target = nnz * k // nparts
        v = int(np.searchsorted(indptr, target, side="left"))
        v = (v + 16) // 32 * 32
        v = max(bounds[-1] + 32, min(v, (V - 32 * (nparts - k)) // 32 * 32))
        bounds.append(v)
    bounds.append(V)
    return bounds


def slab_of(case, rank: int, world: int):
    """(rows slice, D slab CSR, thr, over, under) of `rank`."""
    D = case.D.tocsr()
    b = partition_voxels(D.indptr, world)
    lo, hi = b[rank], b[rank + 1]
    return slice(lo, hi), D[lo:hi], case.thr[lo:hi], case.over[lo:hi], case.under[lo:hi]


class ShardedEvaluator:
    """calcObjGrad over a voxel-sharded D.  `engine` is this rank's slab evaluator with the
    DosePlan interface (set_weights / set_intensities / eval_async(stream, phase) /
    reduce_buffer / fetch_result); `group` a torch.distributed process group (NCCL on GPUs;
    gloo with a host engine in the CPU tests)."""

    def __init__(self, engine, group=None, fused: bool = True):
        """`fused=True` (GPU engines only): exchange over peer memory inside the reduction kernel
        (vmat_comm_attach) instead of an NCCL all-reduce between two kernel phases."""
        import torch.distributed as dist
        self.engine = engine
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.fused = False
        if fused and self.world > 1 and hasattr(engine, "comm_handle"):
            import torch
            rank = dist.get_rank(group)
            mine = torch.tensor(list(engine.comm_handle()), dtype=torch.uint8, device=f"cuda:{engine.device}")
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine, group=group)
            engine.comm_attach(rank, self.world, b"".join(bytes(h.cpu().tolist()) for h in allh))
            dist.barrier(group)
            self.fused = True

    def eval_async(self, stream=None):
        if self.world == 1 or self.fused:
            self.engine.eval_async(stream, 0)
            return
        self.engine.eval_async(stream, 1)                 # local dose, penalty, partial [gb | f]
        buf = self.engine.reduce_buffer()
        self.dist.all_reduce(buf, op=self.dist.ReduceOp.SUM, group=self.group)
        self.engine.eval_async(stream, 2)                 # aperture gradient from the summed gb

    def eval(self, y, stream=None):
        if stream is None and (self.world == 1 or self.fused) and hasattr(self.engine, "eval"):
            return self.engine.eval(y)                    # host buffers in and out, one synchronisation (vmat_eval)
        self.engine.set_intensities(y)
        self.eval_async(stream)
        return self.engine.fetch_result()


def slab_of_device(case, rank: int, world: int):
    """Slab of a DeviceCase (synth.make_device_case): (csr dict of CUDA tensors, thr, over, under)."""
    import torch
    c = case.csr
    if world == 1:
        return c, case.thr, case.over, case.under
    b = partition_voxels(c["vrowptr"].cpu().numpy(), world)
    lo, hi = b[rank], b[rank + 1]
    vr = c["vrowptr"]
    p0, p1 = int(vr[lo].item()), int(vr[hi].item())
    vrowptr = (vr[lo:hi + 1] - vr[lo]).contiguous()
    vcol, vval = c["vcol"][p0:p1].contiguous(), c["vval"][p0:p1].contiguous()
    keep = (c["bcol"] >= lo) & (c["bcol"] < hi)
    B = case.B
    rows_of = torch.repeat_interleave(torch.arange(B, device=keep.device), torch.diff(c["browptr"]))
    browptr = torch.zeros(B + 1, dtype=torch.int64, device=keep.device)
    browptr[1:] = torch.cumsum(torch.bincount(rows_of[keep], minlength=B), 0)
    bcol = (c["bcol"][keep] - lo).to(torch.int32).contiguous()
    bval = c["bval"][keep].contiguous()
    slab = dict(vrowptr=vrowptr, vcol=vcol, vval=vval, browptr=browptr, bcol=bcol, bval=bval)
    return slab, case.thr[lo:hi], case.over[lo:hi], case.under[lo:hi]
