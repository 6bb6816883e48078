"""Greedy k-medoid ordering (kmedoids.py:32-68 of the reference) with the distance pass on
the GPU.  Same three functions; `ba` may be the reference's 1-D list of beam-angle positions
or an (n, d) array of coordinates (the voxel-sampling use of BASELINE.json's north star).

Index semantics.  `kappa` holds ROW indices of the distance matrix, i.e. positions in `ba`.  The
reference takes its candidates from ``listdiff(ba, kappa)`` - the VALUES of `ba` - and uses those
values as row indices too (kmedoids.py:50-53), which only works when the values of a 1-D `ba` are
integers in [0, len(ba)).  That is reproduced: for a 1-D `ba` the candidates are its values, visited in
`ba`'s order (the first best wins, so the order matters), and candidate value i stands for the point
``ba[i]``; values that cannot index the points raise IndexError as they do there.  For an (n, d)
coordinate array (no counterpart in the reference) the candidates are the points 0..n-1 in order.

The reference rebuilds the full pairwise distance matrix for every candidate (O(n^4) for a
whole ordering); here one kernel launch prices all candidates of an insertion step against
the running nearest-medoid distance.  Sums run in the reference's left-to-right order for
n <= 4096 points, so results are bit-identical there.
"""
from __future__ import annotations

import numpy as np

from .plan import kmedoid_costs, KMedoidSession


def _points(ba) -> np.ndarray:
    P = np.asarray(ba, dtype=np.float64)
    if P.ndim == 1:          # the reference embeds 1-D positions as [0, z]  (kmedoids.py:33-34)
        P = np.stack([np.zeros_like(P), P], axis=1)
    return np.ascontiguousarray(P)


def _nearest(P: np.ndarray, kappa) -> np.ndarray:
    """Distance of every point to its nearest medoid (column minimum of D[kappa, :], :38-41)."""
    cur = np.full(P.shape[0], np.inf)
    for k in kappa:
        diff = P - P[k]
        cur = np.minimum(cur, np.sqrt((diff * diff).sum(axis=1)))
    return cur


def distancemin(kappa, ba, device: int = 0):
    """Sum over all points of the distance to the nearest medoid (kmedoids.py:32-42)."""
    P = _points(ba)
    *head, last = list(kappa)
    cost = kmedoid_costs(P, _nearest(P, head), device)
    return cost[last]


def listdiff(a, b):
    b = set(b)
    return [aa for aa in a if aa not in b]


def _sharded_costs(P, cur, device, group, cand=None):
    """Every rank sums over its shard of the points j; the per-rank costs are all-reduced (the voxel
    use of BASELINE.json's north star: the point set is too large for one distance pass).  `cand`:
    candidate coordinates (default: every point)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = P.shape[0]
    j0, j1 = n * rank // world, n * (rank + 1) // world
    part = kmedoid_costs(P[j0:j1], cur[j0:j1], device, candidates=P if cand is None else cand)
    t = torch.from_numpy(part)
    if dist.get_backend(group) == "nccl":
        t = t.to(f"cuda:{device}")
    dist.all_reduce(t, group=group)
    return t.cpu().numpy()


def _candidate_order(kappa, ba, n):
    """None when the candidates are simply the points 0..n-1 in order; else the list of candidate row indices in
    the order the reference visits them (the values of a 1-D `ba`)."""
    a = np.asarray(ba)
    if a.ndim != 1:
        return None
    if a.shape[0] == n and np.array_equal(a, np.arange(n)):
        return None
    vals = [v for v in a.tolist()]
    for v in vals:
        if int(v) != v or not 0 <= int(v) < n:
            raise IndexError(f"kmedoids: value {v} of a 1-D point list cannot index its {n} points "
                             "(the reference uses the values of ba as row indices)")
    return [int(v) for v in vals]


def insertnewelement(kappa, ba, device: int = 0, _state=None, group=None):
    """Append the candidate that minimises distancemin; the first best wins (kmedoids.py:48-60).
    With a torch.distributed `group` the distance pass is sharded over the points."""
    P = _points(ba) if _state is None else _state["P"]
    cur = _nearest(P, kappa) if _state is None else _state["cur"]
    cidx = None if _state is None else _state.get("cand")
    if cidx is not None:        # candidates restricted to a subset of the points (sample_voxels(candidates=...))
        ses = _state.get("session")
        if ses is None:         # this rank's share of the points, the candidates and curmin go to the device once
            if group is None:
                j0, j1 = 0, P.shape[0]
            else:
                import torch.distributed as dist
                rank, world = dist.get_rank(group), dist.get_world_size(group)
                j0, j1 = P.shape[0] * rank // world, P.shape[0] * (rank + 1) // world
            ses = _state["session"] = KMedoidSession(P[j0:j1], cur[j0:j1], _state["Pc"], device)
        if group is None:
            cost = ses.costs()
        else:
            import torch.distributed as dist
            t = ses.costs_device()
            dist.all_reduce(t, group=group)
            cost = t.cpu().numpy()
        cost = np.where(_state["cand_taken"], np.inf, cost)
        k = int(np.argmin(cost))
        _state["cand_taken"][k] = True
        best = int(cidx[k])
        kappa.append(best)
        ses.insert(P[best])     # curmin lives on the device (same operation order as _nearest: bit-identical)
        return kappa
    cost = kmedoid_costs(P, cur, device) if group is None else _sharded_costs(P, cur, device, group)
    order = _candidate_order(kappa, ba, P.shape[0]) if _state is None else _state["order"]
    if order is None:
        taken = np.zeros(P.shape[0], dtype=bool)
        taken[list(kappa)] = True
        cost = np.where(taken, np.inf, cost)
        best = int(np.argmin(cost))          # first minimum == strict "<" scan in candidate order
    else:                                    # 1-D ba that is not 0..n-1: candidates are its values, in its order
        in_kappa = set(kappa)
        cands = [c for c in order if c not in in_kappa]
        best = cands[int(np.argmin(cost[cands]))]
    kappa.append(best)
    if _state is not None:
        diff = P - P[best]
        _state["cur"] = np.minimum(cur, np.sqrt((diff * diff).sum(axis=1)))
    return kappa


def givemewholelist(kappa, ba, device: int = 0, group=None, limit=None, candidates=None):
    """Order all points by greedy k-medoid insertion (kmedoids.py:62-68); `limit` stops after that
    many medoids (voxel subsampling wants the first k, not the whole ordering).  `candidates` (point
    indices, an extension for large voxel sets): only these points may become medoids; the cost of a
    candidate is still summed over ALL points."""
    kappa = list(kappa)
    P = _points(ba)
    state = {"P": P, "cur": _nearest(P, kappa), "order": _candidate_order(kappa, ba, P.shape[0])}
    if candidates is not None:
        cidx = np.asarray(candidates, dtype=np.int64)
        state.update(cand=cidx, Pc=np.ascontiguousarray(P[cidx]), cand_taken=np.isin(cidx, kappa))
        limit = min(limit if limit is not None else len(cidx), len(kappa) + int((~state["cand_taken"]).sum()))
    stop = P.shape[0] if limit is None else min(limit, P.shape[0])
    try:
        while len(kappa) < stop:
            kappa = insertnewelement(kappa, ba, device, state, group)
    finally:
        if state.get("session") is not None:
            state["session"].close()
    return kappa


def sample_voxels(coords, k: int, start: int = 0, device: int = 0, group=None, candidates=None):
    """The first k medoids of the greedy insertion over voxel coordinates (n, 3), starting from voxel
    `start`: a spatially even voxel subsample.  With a torch.distributed `group` every rank prices the
    candidates against its share of the voxels and the costs are all-reduced.  `candidates`: indices of the
    voxels that may be picked (default all; with a million voxels an insertion step over all candidates is
    10^12 distances)."""
    return givemewholelist([int(start)], np.asarray(coords, dtype=np.float64), device=device, group=group, limit=k,
                           candidates=candidates)
