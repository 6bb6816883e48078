"""CORT-format ingest: the on-disk data set the reference's scripts read at start-up
(greedyVMATsampling.py:300-576; formats in SURVEY.md Appendix B) -> one planning case ready
for ``VMATlibrary.bind``.

What the reference does with Python loops over voxels and beamlets is done here with array
operations, and the per-control-point matrices end up as ONE voxel-major matrix D (V x B) in
small voxel space - the form the GPU plan is built from.  Quirks kept on purpose:

* ``pointtoAngle`` covers every angle of range(gastart, gaend, gastep), also those without
  data files, while control points are numbered over the angles that do have files (:68,413-418);
* structures get their single-structure voxel sets in REVERSED priority order, each taking what
  the earlier ones left over (:392-402), and the objective table is reordered by that same
  reversed list (:556);
* beamlets whose (x, y) is not in the intersection of all control points' grids are cut (:513-527).
"""
from __future__ import annotations

import glob
import os
from dataclasses import dataclass, field
from typing import List

import numpy as np
import scipy.io as sio
import scipy.sparse as sp


@dataclass
class Structure:
    """What the reference keeps per structure in ``region`` (:74-91)."""
    index: int
    indices: np.ndarray          # voxels (small space) assigned to this structure only
    fullIndices: np.ndarray      # all voxels (small space) of the structure
    target: bool

    @property
    def sizeInVoxels(self) -> int:
        return len(self.indices)


@dataclass
class LoadedCase:
    V: int
    nb: int
    M: int
    N: int
    beam_ptr: np.ndarray
    angles: List[int]                      # data.pointtoAngle
    gantry: List[int]                      # ga: angles that have data files
    xinter: np.ndarray
    yinter: np.ndarray
    xdirection: List[np.ndarray]
    ydirection: List[np.ndarray]
    D: sp.csr_matrix                       # V x B, float64
    thr: np.ndarray
    over: np.ndarray
    under: np.ndarray
    regions: List[Structure]
    maskValue: np.ndarray
    fullMaskValue: np.ndarray
    voxelAssignment: np.ndarray            # small -> big voxel space
    originalVoxels: np.ndarray             # big -> small (NaN outside every structure)
    names: List[str]
    functionData: np.ndarray               # 3 x numstructs after reordering and size normalisation
    numstructs: int = 0
    numtargets: int = 0
    numoars: int = 0
    regionIndices: List[str] = field(default_factory=list)
    targets: List[str] = field(default_factory=list)
    oars: List[str] = field(default_factory=list)
    numX: int = 0                          # beamlets before the grid cut (sum of numBeamlets)
    totaldijs: int = 0

    @property
    def B(self) -> int:
        return int(self.beam_ptr[-1])

    def Dlist(self):
        Dc = self.D.tocsc()
        return [Dc[:, int(self.beam_ptr[i]):int(self.beam_ptr[i + 1])].T.tocsr() for i in range(self.nb)]


def read_ct_voxel_info(readfolder: str) -> dict:
    """x, y, z voxel counts: the last token of the first three lines (:300-306)."""
    with open(os.path.join(readfolder, "CTVOXEL_INFO.txt")) as f:
        counts = [int(f.readline().split()[-1]) for _ in range(3)]
    return dict(x=counts[0], y=counts[1], z=counts[2])


def _loadmat_sparse(path: str) -> dict:
    try:
        return sio.loadmat(path, spmatrix=True)       # newer scipy: keep the sparse-matrix return type explicit
    except TypeError:
        return sio.loadmat(path)


def _flat_tokens(path: str) -> List[str]:
    with open(path) as f:
        return [tok for line in f for tok in line.rstrip("\n").split("\t")]


def load_cort(readfolder: str, readfolderD: str, structurefile: str, objfile: str, priority,
              gastart: int = 0, gaend: int = 356, gastep: int = 10, algfile: str = None) -> LoadedCase:
    """`priority` is 0-based, least to most important (the reference subtracts 1 at :48)."""
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(readfolder, "*VOILIST.mat")))
    nstruct_files = len(names)
    dims = read_ct_voxel_info(readfolder)
    nbig = dims["x"] * dims["y"] * dims["z"]
    vorg = [np.asarray(sio.loadmat(os.path.join(readfolder, n))["v"]).ravel().astype(np.int64) - 1 for n in names]

    # big <-> small voxel space (:319-354)
    inside = np.zeros(nbig, dtype=bool)
    for v in vorg:
        inside[v] = True
    small_to_big = np.flatnonzero(inside)
    nvox = len(small_to_big)
    big_to_small = np.full(nbig, np.nan)
    big_to_small[small_to_big] = np.arange(nvox)

    # structure map (:356-367)
    tok = _flat_tokens(structurefile)
    numstructs, numtargets, numoars = int(tok[2]), int(tok[3]), int(tok[4])
    region_ids = tok[5:5 + numstructs]
    targets = tok[5 + numstructs:5 + 2 * numstructs]
    oars = tok[5 + 2 * numstructs:5 + 3 * numstructs]

    # mask values (:370-384): bit s of the full mask = voxel belongs to structure s; the single
    # mask keeps the most important structure (priority is least -> most important)
    prio = [int(p) for p in priority]
    small = [big_to_small[v].astype(np.int64) for v in vorg]
    mask_full = np.zeros(nvox)
    mask_single = np.zeros(nvox)
    for s in prio[:nstruct_files]:
        mask_full[small[s]] += 2 ** s
        mask_single[small[s]] = 2 ** s

    # regions in reversed priority, each taking what is still unassigned (:392-402)
    rprio = prio[::-1]
    free = np.ones(nvox, dtype=bool)
    regions = []
    for s in rprio:
        full = small[s]
        own = np.unique(full[free[full]])
        regions.append(Structure(s, own, full, str(s) in targets))
        free[own] = False

    # control points that have both files (:411-418), beamlet grids and their intersection (:444-468)
    gantry = [g for g in range(gastart, gaend, gastep)
              if os.path.isfile(os.path.join(readfolderD, f"Gantry{g}_Couch0_D.mat"))
              and os.path.isfile(os.path.join(readfolder, f"Gantry{g}_Couch0_BEAMINFO.mat"))]
    nb = len(gantry)
    xdir, ydir, nbl, ndij = [], [], [], []
    for g in gantry:
        info = sio.loadmat(os.path.join(readfolder, f"Gantry{g}_Couch0_BEAMINFO.mat"))
        nbl.append(int(np.asarray(info["numBeamlets"]).ravel()[0]))
        ndij.append(int(np.asarray(info["numberNonZerosDij"]).ravel()[0]))
        xdir.append(np.asarray(info["x"])[0])
        ydir.append(np.asarray(info["y"])[0])
    xinter, yinter = xdir[0], ydir[0]
    for i in range(1, nb):
        xinter = np.intersect1d(xinter, xdir[i])
        yinter = np.intersect1d(yinter, ydir[i])

    # dose matrices: big -> small voxel rows, beamlets outside the common grid cut (:487-527)
    blocks = []
    for i, g in enumerate(gantry):
        Dbig = sp.csc_matrix(_loadmat_sparse(os.path.join(readfolderD, f"Gantry{g}_Couch0_D.mat"))["D"], dtype=np.float64)
        keep = np.flatnonzero(np.isin(xdir[i], xinter) & np.isin(ydir[i], yinter))
        blocks.append(Dbig[small_to_big, :][:, keep])
        xdir[i] = xdir[i][keep]
        ydir[i] = ydir[i][keep]
    D = sp.hstack(blocks, format="csr", dtype=np.float64)
    D.sort_indices()
    beam_ptr = np.zeros(nb + 1, dtype=np.int32)
    beam_ptr[1:] = np.cumsum([b.shape[1] for b in blocks])

    # objective table -> per-voxel penalty tables (:534-576)
    fd = np.array([float(t) for t in _flat_tokens(objfile)[3:]]).reshape(3, numstructs)[:, rprio]
    thr, over, under = np.zeros(nvox), np.zeros(nvox), np.zeros(nvox)
    for k, reg in enumerate(regions):
        if reg.sizeInVoxels > 0:
            fd[1, k] = fd[1, k] * 1 / reg.sizeInVoxels
            fd[2, k] = fd[2, k] * 1 / reg.sizeInVoxels
            thr[reg.indices] = fd[0, k]
            over[reg.indices] = fd[1, k]
            under[reg.indices] = fd[2, k]

    return LoadedCase(V=nvox, nb=nb, M=len(xinter), N=len(yinter), beam_ptr=beam_ptr,
                      angles=list(range(gastart, gaend, gastep)), gantry=gantry, xinter=xinter, yinter=yinter,
                      xdirection=xdir, ydirection=ydir, D=D, thr=thr, over=over, under=under, regions=regions,
                      maskValue=mask_single, fullMaskValue=mask_full, voxelAssignment=small_to_big.astype(float),
                      originalVoxels=big_to_small, names=names, functionData=fd, numstructs=numstructs,
                      numtargets=numtargets, numoars=numoars, regionIndices=region_ids, targets=targets, oars=oars,
                      numX=int(sum(nbl)), totaldijs=int(sum(ndij)))
