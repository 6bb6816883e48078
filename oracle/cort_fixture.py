"""TEST INFRASTRUCTURE - write a small synthetic data set in the on-disk formats the reference's
loader reads (greedyVMATsampling.py:300-576): CTVOXEL_INFO.txt,
*_VOILIST.mat, structureInputs.txt, Gantry{g}_Couch0_BEAMINFO.mat, Dij/Gantry{g}_Couch0_D.mat,
obj1.txt, algInputsWilmer.txt.  The CORT data itself is not available; this exercises the same
code paths (big -> small voxel space, priority masks, beamlet-grid intersection cut)."""
from __future__ import annotations

import os

import numpy as np
import scipy.io as sio
import scipy.sparse as sp


def write_cort_like(root: str, seed: int = 0, dims=(12, 10, 6), nstruct: int = 5, gantries=(0, 10, 20, 40),
                    grid=(4, 6)):
    """Returns the configuration a loader needs: paths, priority (0-based), gantry range."""
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "Dij"), exist_ok=True)
    nx, ny, nz = dims
    nvox = nx * ny * nz
    with open(os.path.join(root, "CTVOXEL_INFO.txt"), "w") as f:
        f.write(f"xdim voxels {nx}\nydim voxels {ny}\nzdim voxels {nz}\nspacing 0.3\n")
    # structures: overlapping blobs covering ~60 % of the big voxel space (the rest is "air")
    centre = rng.random((nstruct, 3)) * np.array(dims)
    v = np.arange(nvox)
    xyz = np.stack([v % nx, (v // nx) % ny, v // (nx * ny)], axis=1).astype(float)
    in_any = np.zeros(nvox, dtype=bool)
    names = []
    for s in range(nstruct):
        dist = np.sqrt(((xyz - centre[s]) ** 2).sum(axis=1))
        members = np.flatnonzero(dist < np.quantile(dist, 0.18 + 0.06 * s))
        in_any[members] = True
        name = f"S{s:02d}_{'PTV' if s < 2 else 'OAR'}_VOILIST.mat"
        names.append(name)
        sio.savemat(os.path.join(root, name), {"v": (members + 1).reshape(-1, 1).astype(np.int32)})
    covered = np.flatnonzero(in_any)
    # structure map file: 2 header tokens, then numstructs numtargets numoars, region ids, targets, oars
    region_ids = [str(s) for s in range(nstruct)]
    targets = [str(s) if s < 2 else "-1" for s in range(nstruct)]
    oars = [str(s) if s >= 2 else "-1" for s in range(nstruct)]
    with open(os.path.join(root, "structureInputs.txt"), "w") as f:
        f.write("\t".join(["structures", "v1", str(nstruct), "2", str(nstruct - 2)] + region_ids + targets + oars) + "\n")
    # objective file: 3 header tokens then thresholds, over weights, under weights
    thr = [1.0 if s < 2 else round(0.2 + 0.1 * s, 2) for s in range(nstruct)]
    over = [100.0 if s < 2 else 5.0 + s for s in range(nstruct)]
    under = [100.0 if s < 2 else 0.0 for s in range(nstruct)]
    with open(os.path.join(root, "obj1.txt"), "w") as f:
        f.write("\t".join(["obj", "quad", "v1"] + [repr(float(t)) for t in thr + over + under]) + "\n")
    with open(os.path.join(root, "algInputsWilmer.txt"), "w") as f:
        f.write("alg\tnone\n")
    # beams: every gantry has the full M x N grid plus a few extra beamlets outside the common grid
    # (the loader must cut those away) and a few grid beamlets missing at the row ends (ragged rows)
    M, N = grid
    xs = np.arange(M) * 0.5 - 0.25 * (M - 1)
    ys = np.arange(N) * 0.5 - 0.25 * (N - 1)
    for gi, g in enumerate(gantries):
        bx, by = [], []
        for m in range(M):
            lo = 1 if (m + gi) % 3 == 0 and m not in (0, M - 1) else 0
            hi = N - 1 if (m + gi) % 4 == 1 and m not in (0, M - 1) else N
            for n in range(lo, hi):
                bx.append(xs[m]); by.append(ys[n])
        # rows 0 and M-1 are complete in every beam, so the intersections keep all x and y values
        extra = 2 + gi % 2
        for e in range(extra):                         # beamlets off the common grid
            bx.append(xs[-1] + 0.5 * (e + 1) + 0.01 * (gi + 1)); by.append(ys[e % N])   # x differs per beam: cut by the intersection
        bx, by = np.array(bx), np.array(by)
        nbl = len(bx)
        dens = 0.08
        nnz = int(dens * len(covered) * nbl)
        rows = covered[rng.integers(0, len(covered), size=nnz)]
        cols = rng.integers(0, nbl, size=nnz)
        vals = np.float32(0.05 + rng.random(nnz)).astype(np.float64)
        D = sp.csc_matrix((vals, (rows, cols)), shape=(nvox, nbl))
        D.sum_duplicates()
        D.data = np.float32(D.data).astype(np.float64)
        sio.savemat(os.path.join(root, f"Gantry{g}_Couch0_BEAMINFO.mat"),
                    {"numBeamlets": float(nbl), "numberNonZerosDij": float(D.nnz), "x": bx.reshape(1, -1), "y": by.reshape(1, -1)})
        sio.savemat(os.path.join(root, "Dij", f"Gantry{g}_Couch0_D.mat"), {"D": D})
    priority = list(rng.permutation(nstruct))
    return dict(readfolder=root + os.sep, readfolderD=os.path.join(root, "Dij") + os.sep,
                structurefile=os.path.join(root, "structureInputs.txt"), objfile=os.path.join(root, "obj1.txt"),
                algfile=os.path.join(root, "algInputsWilmer.txt"), priority=[int(p) for p in priority],
                gastart=0, gaend=max(gantries) + 1, gastep=10)
