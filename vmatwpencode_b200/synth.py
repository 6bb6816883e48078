"""Synthetic VMAT cases (phantom + dose-deposition matrix D + penalty tables).

The reference ships no data and no generator: it reads the CORT head-and-neck
``.mat`` files from the author's disk (greedyVMATsampling.py:308-576).  This
module builds cases of the same *shape* so that the hot path can be exercised
and measured: a voxel phantom, ``nb`` control points on an arc, an M x N
beamlet grid per control point (x-major beamlet order, the order
``fvalidbeamlets`` assumes, greedyVMATsampling.py:589-602), a sparse D with a
pencil-beam ("geometric") or uniform-random sparsity pattern, and the three
per-voxel penalty tables ``quadHelperThresh/Over/Under``
(greedyVMATsampling.py:554-576: threshold, over-weight/nS, under-weight/nS).

Everything is a pure function of its arguments (no RNG state): values come
from integer hashes so a case is identical wherever it is generated.  D values
are rounded to the storage dtype (fp32 or bf16), so the float64 CPU oracle and
the GPU kernels consume bit-identical numbers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import scipy.sparse as sp
import torch

# name -> (V, nb, M, N, density, seed); BASELINE.json configs / SURVEY.md 8(d)
CONFIGS = {
    "cfg1": dict(V=20_000, nb=36, M=7, N=8, density=0.01, seed=1001, gastep=10),
    "cfg2": dict(V=300_000, nb=180, M=7, N=8, density=0.01, seed=1002, gastep=2),
    "cfg3": dict(V=1_000_000, nb=180, M=9, N=9, density=0.01, seed=1003, gastep=2),
    "cfg4": dict(V=4_000_000, nb=360, M=9, N=9, density=0.01, seed=1004, gastep=1),
}


@dataclass
class Case:
    """One synthetic planning case, in the vocabulary of the reference."""
    V: int                      # data.numvoxels
    nb: int                     # data.numbeams (control points)
    M: int                      # leaf rows  = len(data.xinter)
    N: int                      # beamlets per row = len(data.yinter)
    beam_ptr: np.ndarray        # int32[nb+1], beamlet offset of each control point
    angles: List[int]           # data.pointtoAngle
    xinter: np.ndarray          # f64[M]
    yinter: np.ndarray          # f64[N]
    xdirection: List[np.ndarray]  # per beam f64[B_i] (x-major order)
    ydirection: List[np.ndarray]
    D: sp.csr_matrix            # V x B float64, values representable in `val_dtype`
    thr: np.ndarray             # quadHelperThresh  f64[V]
    over: np.ndarray            # quadHelperOver    f64[V]
    under: np.ndarray           # quadHelperUnder   f64[V]
    struct_id: np.ndarray       # uint8[V]
    val_dtype: str = "f32"
    meta: dict = field(default_factory=dict)

    @property
    def B(self) -> int:
        return int(self.beam_ptr[-1])

    @property
    def nnz(self) -> int:
        return int(self.D.nnz)

    def Dlist(self) -> List[sp.csr_matrix]:
        """Per-control-point (B_i x V) CSR matrices, the reference's ``data.Dlist``
        (greedyVMATsampling.py:482,524)."""
        Dc = self.D.tocsc()
        out = []
        for i in range(self.nb):
            lo, hi = int(self.beam_ptr[i]), int(self.beam_ptr[i + 1])
            out.append(Dc[:, lo:hi].T.tocsr())
        return out


def _hash01(a: torch.Tensor, b: torch.Tensor, seed: int) -> torch.Tensor:
    """Deterministic integer hash of (a, b, seed) -> float64 in [0, 1)."""
    h = (a.to(torch.int64) * 2654435761 + b.to(torch.int64) * 40503 + seed * 97 + 12345) & 0xFFFFFFFF
    h = ((h ^ (h >> 15)) * 2246822519) & 0xFFFFFFFF
    h = ((h ^ (h >> 13)) * 3266489917) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    return h.to(torch.float64) / 4294967296.0


def _round_to(vals: torch.Tensor, val_dtype: str) -> torch.Tensor:
    if val_dtype == "f32":
        return vals.to(torch.float32).to(torch.float64)
    if val_dtype == "bf16":
        return vals.to(torch.bfloat16).to(torch.float64)
    if val_dtype == "f64":
        return vals
    raise ValueError(val_dtype)


def _field_halfwidth(frac: float) -> float:
    """Half-width W of the strip |u|<=W whose intersection with the unit disk
    covers `frac` of the [-1,1]^2 square (area 4)."""
    frac = min(frac, math.pi / 4 - 1e-9)
    lo, hi = 0.0, 1.0
    for _ in range(60):
        w = 0.5 * (lo + hi)
        area = 2.0 * (w * math.sqrt(1 - w * w) + math.asin(w))
        if area / 4.0 < frac:
            lo = w
        else:
            hi = w
    return 0.5 * (lo + hi)


def phantom_grid(V: int, M: int):
    """Voxel grid (nx, ny, nz) with nz a multiple of M; the first V cells are used."""
    nz = M * max(1, round((V ** (1.0 / 3.0)) / M))
    nxy = max(1, math.ceil(math.sqrt(V / nz)))
    while nxy * nxy * nz < V:
        nxy += 1
    return nxy, nxy, nz


def voxel_coordinates(V: int, M: int) -> np.ndarray:
    """(V, 3) grid indices of the phantom's voxels (integers, so distances between them are exact in
    float64): the point set of the k-medoid voxel subsampling."""
    nx, ny, _ = phantom_grid(V, M)
    v = np.arange(V, dtype=np.int64)
    return np.stack([v % nx, (v // nx) % ny, v // (nx * ny)], axis=1).astype(np.float64)


def subsample_case(case: "Case", rows) -> "Case":
    """The case restricted to the voxels `rows` (ascending order kept): their rows of D and their
    penalty entries, the weights rescaled per structure by (voxels in structure) / (sampled voxels in
    structure) so that objective values stay comparable with the full case.  (The reference samples
    control points, not voxels; this is the voxel subsampling BASELINE.json's third configuration names.)"""
    rows = np.sort(np.asarray(rows, dtype=np.int64))
    sid = np.asarray(case.struct_id)
    full = np.bincount(sid, minlength=256).astype(np.float64)
    kept = np.bincount(sid[rows], minlength=256).astype(np.float64)
    scale = np.divide(full, kept, out=np.ones(256), where=kept > 0)[sid[rows]]
    D = case.D.tocsr()[rows]
    return Case(V=len(rows), nb=case.nb, M=case.M, N=case.N, beam_ptr=case.beam_ptr, angles=list(case.angles),
                xinter=case.xinter, yinter=case.yinter, xdirection=case.xdirection, ydirection=case.ydirection,
                D=D, thr=case.thr[rows].copy(), over=case.over[rows] * scale, under=case.under[rows] * scale,
                struct_id=sid[rows].copy(), val_dtype=case.val_dtype, meta=dict(case.meta, subsampled_from=case.V))


def subsample_device_case(dcase: "DeviceCase", rows, rich_structures: bool = False) -> "Case":
    """subsample_case for a device-generated case: the rows of D are gathered on the GPU and brought to
    the host as a scipy matrix; beamlet geometry is the full M x N rectangle per control point the device
    generator uses."""
    rows = np.sort(np.asarray(rows, dtype=np.int64))
    c = dcase.csr
    dev = c["vval"].device
    r = torch.as_tensor(rows, device=dev)
    lo, hi = c["vrowptr"][r], c["vrowptr"][r + 1]
    n = hi - lo
    pos = torch.arange(int(n.sum().item()), device=dev) - torch.repeat_interleave(torch.cumsum(n, 0) - n, n) + \
        torch.repeat_interleave(lo, n)
    indptr = np.concatenate([[0], np.cumsum(n.cpu().numpy())])
    D = sp.csr_matrix((c["vval"][pos].double().cpu().numpy(), c["vcol"][pos].cpu().numpy(), indptr),
                      shape=(len(rows), dcase.B))
    sid, _ = structures(dcase.V, dcase.M, device=dev, rich=rich_structures)     # same device as the case: same memberships
    sid = sid.cpu().numpy()
    full = np.bincount(sid, minlength=256).astype(np.float64)
    kept = np.bincount(sid[rows], minlength=256).astype(np.float64)
    scale = np.divide(full, kept, out=np.ones(256), where=kept > 0)[sid[rows]]
    M, N = dcase.M, dcase.N
    xinter = np.arange(M, dtype=np.float64) * 0.5 - 0.25 * (M - 1)
    yinter = np.arange(N, dtype=np.float64) * 0.5 - 0.25 * (N - 1)
    xd = [np.repeat(xinter, N) for _ in range(dcase.nb)]
    yd = [np.tile(yinter, M) for _ in range(dcase.nb)]
    return Case(V=len(rows), nb=dcase.nb, M=M, N=N, beam_ptr=dcase.beam_ptr, angles=list(dcase.angles),
                xinter=xinter, yinter=yinter, xdirection=xd, ydirection=yd, D=D,
                thr=dcase.thr[rows].copy(), over=dcase.over[rows] * scale, under=dcase.under[rows] * scale,
                struct_id=sid[rows].copy(), val_dtype="f32", meta=dict(dcase.meta, subsampled_from=dcase.V))


def structures(V: int, M: int, device="cpu", rich: bool = False):
    """struct_id[V] plus the per-structure (threshold, over, under) table.

    Default: PTV 10 % (t=1, over=under=100), OAR 30 % (t=0.3, over=10),
    body 60 % (t=0.5, over=1); weights are divided by structure size like
    greedyVMATsampling.py:558-561.  `rich` adds a second PTV and more OARs.
    """
    nx, ny, nz = phantom_grid(V, M)
    v = torch.arange(V, device=device, dtype=torch.int64)
    x = ((v % nx).to(torch.float64) + 0.5) / nx * 2 - 1
    y = (((v // nx) % ny).to(torch.float64) + 0.5) / ny * 2 - 1
    z = ((v // (nx * ny)).to(torch.float64) + 0.5) / nz * 2 - 1
    sid = torch.full((V,), 255, dtype=torch.uint8, device=device)
    if not rich:
        specs = [  # (centre, fraction, thr, over, under)
            ((0.0, 0.0, 0.0), 0.10, 1.0, 100.0, 100.0),
            ((0.35, -0.25, 0.1), 0.30, 0.3, 10.0, 0.0),
        ]
        rest = (0.5, 1.0, 0.0)
    else:
        specs = [
            ((0.0, 0.0, 0.0), 0.06, 1.0, 100.0, 100.0),
            ((-0.3, 0.3, 0.3), 0.04, 0.9, 80.0, 80.0),
            ((0.4, -0.3, 0.0), 0.08, 0.3, 10.0, 0.0),
            ((-0.4, -0.4, -0.2), 0.06, 0.25, 12.0, 0.0),
            ((0.5, 0.4, 0.4), 0.05, 0.35, 8.0, 0.0),
            ((0.0, 0.6, -0.5), 0.05, 0.2, 15.0, 0.0),
            ((0.0, -0.6, 0.5), 0.04, 0.4, 6.0, 0.0),
            ((-0.6, 0.0, 0.0), 0.04, 0.45, 5.0, 0.0),
        ]
        rest = (0.5, 1.0, 0.0)
    table = []
    for s, (c, frac, t, ow, uw) in enumerate(specs):
        free = sid == 255
        dist = (x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2
        dist = torch.where(free, dist, torch.full_like(dist, float("inf")))
        k = max(1, min(int(round(frac * V)), int(free.sum().item())))
        kth = torch.kthvalue(dist, k).values
        pick = free & (dist <= kth)
        sid[pick] = s
        table.append((t, ow, uw))
    sid[sid == 255] = len(specs)
    table.append(rest)
    return sid, table


def penalty_tables(sid: torch.Tensor, table):
    """Per-voxel quadHelperThresh/Over/Under (greedyVMATsampling.py:564-576)."""
    V = sid.numel()
    thr = torch.zeros(V, dtype=torch.float64, device=sid.device)
    over = torch.zeros_like(thr)
    under = torch.zeros_like(thr)
    for s, (t, ow, uw) in enumerate(table):
        m = sid == s
        n = int(m.sum().item())
        if n == 0:
            continue
        thr[m] = t
        over[m] = ow * 1 / n
        under[m] = uw * 1 / n
    return thr, over, under


def beam_entries(i: int, theta_deg: float, V: int, M: int, N: int, frac_in_field: float,
                 seed: int, device="cpu", row_lo=None, row_hi=None):
    """COO entries (voxel, local beamlet, value) of control point i (geometric model).

    A voxel at lateral offset u = -x sin(t) + y cos(t) inside the field strip is
    hit by the beamlet of its z band (leaf row m) and lateral bin (n), plus the
    nearer neighbouring bin at 30 %; dose falls off exponentially with depth.
    """
    nx, ny, nz = phantom_grid(V, M)
    v = torch.arange(V, device=device, dtype=torch.int64)
    x = ((v % nx).to(torch.float64) + 0.5) / nx * 2 - 1
    y = (((v // nx) % ny).to(torch.float64) + 0.5) / ny * 2 - 1
    zi = v // (nx * ny)
    th = math.radians(theta_deg)
    u = -x * math.sin(th) + y * math.cos(th)
    s = x * math.cos(th) + y * math.sin(th)
    W = _field_halfwidth(frac_in_field)
    infield = (u.abs() < W) & (x * x + y * y <= 1.0)
    m = torch.clamp(zi * M // nz, max=M - 1)
    t = (u + W) / (2 * W) * N
    n = torch.clamp(t.floor().to(torch.int64), 0, N - 1)
    fracpos = t - n.to(torch.float64)
    nbr = torch.where(fracpos >= 0.5, n + 1, n - 1)
    vsel = v[infield]
    m_s, n_s, nbr_s, s_s = m[infield], n[infield], nbr[infield], s[infield]
    depth = torch.exp(-0.6 * (s_s + 1.2))
    main_v = depth * (0.8 + 0.4 * _hash01(vsel, n_s + 64 * i, seed))
    nb_ok = (nbr_s >= 0) & (nbr_s < N)
    side_v = 0.3 * depth * (0.8 + 0.4 * _hash01(vsel, nbr_s + 64 * i + 7919, seed))
    rows = torch.cat([vsel, vsel[nb_ok]])
    ncol = torch.cat([n_s, nbr_s[nb_ok]])
    mrow = torch.cat([m_s, m_s[nb_ok]])
    vals = torch.cat([main_v, side_v[nb_ok]])
    if row_lo is not None:  # ragged rows: keep only beamlets n in [row_lo[m], row_hi[m])
        lo = row_lo[mrow]
        hi = row_hi[mrow]
        keep = (ncol >= lo) & (ncol < hi)
        rows, ncol, mrow, vals = rows[keep], ncol[keep], mrow[keep], vals[keep]
    return rows, mrow, ncol, vals


def make_case(V: int, nb: int, M: int, N: int, density: float = 0.01, seed: int = 0,
              model: str = "geometric", val_dtype: str = "f32", gastep: Optional[int] = None,
              ragged: bool = False, rich_structures: bool = False, device: str = "cpu") -> Case:
    """Build a synthetic case.  `model` is "geometric" or "uniform"."""
    if gastep is None:
        gastep = max(1, 360 // nb)
    angles = [gastep * i for i in range(nb)]
    xinter = np.arange(M, dtype=np.float64) * 0.5 - 0.25 * (M - 1)
    yinter = np.arange(N, dtype=np.float64) * 0.5 - 0.25 * (N - 1)

    # per beam, per row: contiguous run of valid y indices [lo, hi)
    row_lo = np.zeros((nb, M), dtype=np.int64)
    row_hi = np.full((nb, M), N, dtype=np.int64)
    if ragged:
        ii = torch.arange(nb * M, dtype=torch.int64)
        h = _hash01(ii, ii * 0 + 3, seed).numpy().reshape(nb, M)
        g = _hash01(ii, ii * 0 + 5, seed).numpy().reshape(nb, M)
        row_lo = np.minimum((h * 3).astype(np.int64) % 3, max(0, N - 2))
        row_hi = np.maximum(N - ((g * 3).astype(np.int64) % 3), row_lo + 2)
        row_hi = np.minimum(row_hi, N)
        # every y index and every x index must survive in some beam's rows so that
        # xinter/yinter (intersections over beams) stay full; force beam rows to
        # share the y extremes at least once per beam
        row_lo[:, 0] = 0
        row_hi[:, 0] = N
    blen = (row_hi - row_lo)
    beam_ptr = np.zeros(nb + 1, dtype=np.int32)
    beam_ptr[1:] = np.cumsum(blen.sum(axis=1))
    rowstart = np.zeros((nb, M), dtype=np.int64)   # local offset of row m in beam i
    rowstart[:, 1:] = np.cumsum(blen, axis=1)[:, :-1]
    xdirection, ydirection = [], []
    for i in range(nb):
        xd, yd = [], []
        for m in range(M):
            for n in range(row_lo[i, m], row_hi[i, m]):
                xd.append(xinter[m])
                yd.append(yinter[n])
        xdirection.append(np.array(xd))
        ydirection.append(np.array(yd))
    B = int(beam_ptr[-1])

    if model == "geometric":
        per_hit = 2.0 - 1.0 / (2 * N)  # main bin + the nearer neighbour when it exists
        frac = density * M * N / per_hit
        frac = min(frac, 0.7)
        r_all, c_all, v_all = [], [], []
        t_lo = torch.as_tensor(row_lo, device=device)
        t_hi = torch.as_tensor(row_hi, device=device)
        t_rs = torch.as_tensor(rowstart, device=device)
        for i in range(nb):
            rows, mrow, ncol, vals = beam_entries(i, angles[i], V, M, N, frac, seed, device,
                                                  t_lo[i] if ragged else None, t_hi[i] if ragged else None)
            col = int(beam_ptr[i]) + t_rs[i][mrow] + (ncol - t_lo[i][mrow])
            r_all.append(rows)
            c_all.append(col)
            v_all.append(_round_to(vals, val_dtype))
        rows = torch.cat(r_all).cpu().numpy()
        cols = torch.cat(c_all).cpu().numpy()
        vals = torch.cat(v_all).cpu().numpy()
    elif model == "uniform":
        nnz_target = int(round(density * V * B))
        k = torch.arange(nnz_target, dtype=torch.int64, device=device)
        rows_t = (_hash01(k, k * 0 + 11, seed) * V).to(torch.int64).clamp_(max=V - 1)
        cols_t = (_hash01(k, k * 0 + 13, seed + 1) * B).to(torch.int64).clamp_(max=B - 1)
        key = torch.unique(rows_t * B + cols_t)
        rows_t, cols_t = key // B, key % B
        vals_t = _round_to(0.05 + _hash01(rows_t, cols_t, seed + 2), val_dtype)
        rows, cols, vals = rows_t.cpu().numpy(), cols_t.cpu().numpy(), vals_t.cpu().numpy()
    else:
        raise ValueError(model)

    D = sp.csr_matrix((vals, (rows, cols)), shape=(V, B), dtype=np.float64)
    D.sum_duplicates()
    D.sort_indices()
    if val_dtype != "f64":   # duplicate sums (none expected) must stay representable
        D.data = _round_to(torch.from_numpy(D.data), val_dtype).numpy()

    sid, table = structures(V, M, device=device, rich=rich_structures)
    thr, over, under = penalty_tables(sid, table)
    return Case(V=V, nb=nb, M=M, N=N, beam_ptr=beam_ptr, angles=angles, xinter=xinter, yinter=yinter,
                xdirection=xdirection, ydirection=ydirection, D=D,
                thr=thr.cpu().numpy(), over=over.cpu().numpy(), under=under.cpu().numpy(),
                struct_id=sid.cpu().numpy(), val_dtype=val_dtype,
                meta=dict(model=model, density_target=density, density=D.nnz / (V * max(B, 1)),
                          seed=seed, gastep=gastep, ragged=ragged,
                          row_lo=row_lo, row_hi=row_hi))


def make_config(name: str, scale: float = 1.0, **overrides) -> Case:
    """One of BASELINE.json's configurations (optionally with V scaled down)."""
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    cfg["V"] = max(64, int(cfg["V"] * scale))
    return make_case(**cfg)


def case_to_arrays(case: Case, prefix: str = "case_") -> dict:
    """Flatten a Case into plain arrays (for .npz fixtures)."""
    D = case.D.tocsr()
    out = dict(V=case.V, nb=case.nb, M=case.M, N=case.N, beam_ptr=case.beam_ptr,
               angles=np.asarray(case.angles), xinter=case.xinter, yinter=case.yinter,
               xdir=np.concatenate(case.xdirection), ydir=np.concatenate(case.ydirection),
               indptr=D.indptr.astype(np.int64), indices=D.indices.astype(np.int32), data=D.data,
               thr=case.thr, over=case.over, under=case.under, struct_id=case.struct_id)
    return {prefix + k: np.asarray(v) for k, v in out.items()}


def case_from_arrays(z, prefix: str = "case_") -> Case:
    g = lambda k: z[prefix + k]
    V, nb, M, N = int(g("V")), int(g("nb")), int(g("M")), int(g("N"))
    beam_ptr = np.asarray(g("beam_ptr"), dtype=np.int32)
    xdir, ydir = g("xdir"), g("ydir")
    xd = [np.array(xdir[beam_ptr[i]:beam_ptr[i + 1]]) for i in range(nb)]
    yd = [np.array(ydir[beam_ptr[i]:beam_ptr[i + 1]]) for i in range(nb)]
    D = sp.csr_matrix((g("data"), g("indices"), g("indptr")), shape=(V, int(beam_ptr[-1])))
    return Case(V=V, nb=nb, M=M, N=N, beam_ptr=beam_ptr, angles=[int(a) for a in g("angles")],
                xinter=np.array(g("xinter")), yinter=np.array(g("yinter")), xdirection=xd, ydirection=yd,
                D=D, thr=np.array(g("thr")), over=np.array(g("over")), under=np.array(g("under")),
                struct_id=np.array(g("struct_id")))


@dataclass
class DeviceCase:
    """A synthetic case whose D lives on the GPU only (both CSR orientations as CUDA tensors);
    for configurations too large to build through scipy on the host (configs 3-5)."""
    V: int
    nb: int
    M: int
    N: int
    beam_ptr: np.ndarray
    angles: List[int]
    thr: np.ndarray
    over: np.ndarray
    under: np.ndarray
    csr: dict                   # vrowptr, vcol, vval, browptr, bcol, bval (torch, on device)
    nnz: int
    meta: dict = field(default_factory=dict)

    @property
    def B(self) -> int:
        return int(self.beam_ptr[-1])


def make_device_case(V: int, nb: int, M: int, N: int, density: float = 0.01, seed: int = 0,
                     model: str = "geometric", val_dtype: str = "f32", gastep: Optional[int] = None,
                     rich_structures: bool = False, device: str = "cuda", **_ignored) -> DeviceCase:
    """Same cases as make_case (full M x N beamlet rectangles), generated and sorted on the GPU."""
    if gastep is None:
        gastep = max(1, 360 // nb)
    angles = [gastep * i for i in range(nb)]
    B = nb * M * N
    beam_ptr = (np.arange(nb + 1, dtype=np.int64) * (M * N)).astype(np.int32)
    keys, vals = [], []
    if model == "geometric":
        per_hit = 2.0 - 1.0 / (2 * N)
        frac = min(density * M * N / per_hit, 0.7)
        for i in range(nb):
            rows, mrow, ncol, v = beam_entries(i, angles[i], V, M, N, frac, seed, device)
            col = i * M * N + mrow * N + ncol
            keys.append(rows * B + col)
            vals.append(_round_to(v, val_dtype).to(torch.float32 if val_dtype != "f64" else torch.float64))
        key = torch.cat(keys); val = torch.cat(vals)
        del keys, vals
    elif model == "uniform":
        nnz_target = int(round(density * V * B))
        k = torch.arange(nnz_target, dtype=torch.int64, device=device)
        rows_t = (_hash01(k, k * 0 + 11, seed) * V).to(torch.int64).clamp_(max=V - 1)
        cols_t = (_hash01(k, k * 0 + 13, seed + 1) * B).to(torch.int64).clamp_(max=B - 1)
        key = torch.unique(rows_t * B + cols_t)
        del rows_t, cols_t, k
        val = _round_to(0.05 + _hash01(key // B, key % B, seed + 2), val_dtype).to(
            torch.float32 if val_dtype != "f64" else torch.float64)
    else:
        raise ValueError(model)
    # voxel-major: sort by (row, col)
    key, order = torch.sort(key)
    vval = val[order]
    del order, val
    rows = key // B
    vcol = (key - rows * B).to(torch.int32)
    vrowptr = torch.zeros(V + 1, dtype=torch.int64, device=device)
    vrowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=V), 0)
    # beamlet-major: sort by (col, row)
    key2 = vcol.to(torch.int64) * V + rows
    del key, rows
    key2, order = torch.sort(key2)
    bval = vval[order]
    del order
    cols = key2 // V
    bcol = (key2 - cols * V).to(torch.int32)
    browptr = torch.zeros(B + 1, dtype=torch.int64, device=device)
    browptr[1:] = torch.cumsum(torch.bincount(cols, minlength=B), 0)
    del key2, cols
    sid, table = structures(V, M, device=device, rich=rich_structures)
    thr, over, under = penalty_tables(sid, table)
    nnz = int(vcol.numel())
    return DeviceCase(V=V, nb=nb, M=M, N=N, beam_ptr=beam_ptr, angles=angles,
                      thr=thr.cpu().numpy(), over=over.cpu().numpy(), under=under.cpu().numpy(),
                      csr=dict(vrowptr=vrowptr, vcol=vcol, vval=vval, browptr=browptr, bcol=bcol, bval=bval),
                      nnz=nnz, meta=dict(model=model, density=nnz / (V * B), seed=seed, gastep=gastep))


def make_device_config(name: str, scale: float = 1.0, **overrides) -> DeviceCase:
    cfg = dict(CONFIGS[name])
    cfg.update(overrides)
    cfg["V"] = max(64, int(cfg["V"] * scale))
    return make_device_case(**cfg)
