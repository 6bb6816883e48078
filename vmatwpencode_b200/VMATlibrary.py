"""Drop-in objective / gradient / pricing functions on the B200 path.

The reference's driver scripts (greedyVMATsampling.py and its earlier generations)
define these names at module level and let them close over the singleton ``data``:

    data.calcDose(), data.calcGradientandObjValue()      greedyVMATsampling.py:243-272
    calcObjGrad(x, user_data=None)                        :578-582
    fvalidbeamlets(i, index)                              :589-602
    PPsubroutine(C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim,
                 predec, succ, N, M, thisApertureIndex, bw)              :618-789
    parallelizationPricingProblem(i, C, C2, C3, b, vmax, speedlim, N, M, bw)   :791-822
    PricingProblem(C, C2, C3, b, vmax, speedlim, N, M, bw)               :824-856
    solveRMC(YU)                                          :858-885
    colGen(C, WholeCircle, initialApertures)              :937-1061
    updateOpenAperture(i)                                 :1070-1130

This module provides the same names with the same arguments, return values and error
behaviour (sampling-variant semantics, quirks Q1-Q8 of SURVEY.md included), but the sparse
arithmetic runs in libvmat_b200.so on the GPU: ``bind(case)`` uploads D once, and each
``calcObjGrad`` is three CUDA kernels.  There is no CPU fallback.

Not reproduced: plotting, pickling and the debug prints of the reference.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import random
import sys
import time
import warnings

import numpy as np
from scipy.optimize import minimize

from ._lib import PriceParams, PriceRow, VmatError
from .plan import DosePlan

# module globals the reference scripts also have
data = None
N = None          # beamlets per leaf row  (len(data.yinter), :466)
M = None          # leaf rows              (len(data.xinter), :468)
kappa = []        # candidate ordering used by the WholeCircle branch (:30,959-966)
eliminationThreshold = 10E-3
eliminationPhase = False     # the reference hard-codes False inside colGen (:940); settable here
kappasize = 16
numcores = 8
use_column_cache = False     # True: solveRMC evaluates from cached dense aperture columns (pays off when an
                             # RMP needs many evaluations; with the reference's ftol=1e-1 it needs ~5)
penalizationweights = []     # Kelly complexity measure per selected control point (Laptop variant :484,863)
fast_rmp = True              # solveRMC drives scipy's own L-BFGS-B routine (scipy.optimize._lbfgsb.setulb) with a lean loop:
                             # same routine, same parameters, same iterates as minimize(method="L-BFGS-B"), without the
                             # ~1.5 ms of Python set-up per call and ~0.3 ms per evaluation that minimize() spends around it


class region:
    """Structure bookkeeping (greedyVMATsampling.py:74-91)."""

    def __init__(self, iind, iindi, ifullindi, itarget):
        self.index = iind
        self.sizeInVoxels = len(iindi)
        self.indices = iindi
        self.fullIndices = ifullindi
        self.target = itarget


class apertureList:
    """Control-point locations with their gantry angles, kept ascending
    (greedyVMATsampling.py:97-140).  loc and angle are sorted separately, as there."""

    def __init__(self):
        self.loc = []
        self.angle = []

    def insertAngle(self, i, aperangle):
        self.angle.append(aperangle)
        self.loc.append(i)
        self.loc.sort()
        self.angle.sort()

    def removeIndex(self, index):
        k = self.loc.index(index)
        del self.loc[k]
        del self.angle[k]

    def removeAngle(self, tangl):
        k = self.angle.index(tangl)
        del self.loc[k]
        del self.angle[k]

    def __call__(self, index):
        return self.angle[self.loc.index(index)]

    def len(self):
        return len(self.loc)

    def isEmpty(self):
        return not self.loc


class _LazyDZdK:
    """Stand-in for the reference's dense V x numbeams ``dZdK`` (:246,251).  It is never
    formed on the GPU path (11.5 GB at 4M voxels x 360 control points): aperturegradient comes
    from the beamlet gradient instead, and asking for the dense matrix is an error."""

    def __init__(self, owner):
        self._owner = owner

    def toarray(self):
        raise VmatError("dZdK is not materialised on the GPU path; aperturegradient is computed "
                        "as sum_j w_grad[j] * (D^T vg)[j]")


class vmat_class:
    """The reference's global state object (greedyVMATsampling.py:145-276) with the two
    evaluation methods backed by a DosePlan."""

    def __init__(self):
        self.numX = 0
        self.numvoxels = 0
        self.numstructs = 0
        self.numoars = 0
        self.numtargets = 0
        self.numbeams = 0
        self.totaldijs = 0
        self.objectiveValue = float("inf")
        self.beamletsPerBeam = []
        self.dijsPerBeam = []
        self.maskValue = []
        self.fullMaskValue = []
        self.regionIndices = []
        self.targets = []
        self.oars = []
        self.regions = []
        self.functionData = []
        self.notinC = apertureList()
        self.caligraphicC = apertureList()
        self.currentIntensities = np.empty(1)
        self.xinter = []
        self.yinter = []
        self.xdirection = []
        self.ydirection = []
        self.llist = []
        self.rlist = []
        self.rmpres = []
        self.aperturegradient = []
        self.openApertureMaps = []
        self.diagmakers = []
        self.strengths = []
        self.dZdK = 0.0
        self.pointtoAngle = []
        self.Dlist = []
        self.entryCounter = 0
        self.listIndexofAperturesRemovedEachStep = []
        self.history = []
        # GPU side
        self.plan = None
        self.beam_ptr = None
        self._uploaded = {}
        self._dose = None
        self._vg = None
        self._gb = None
        self._fg = None
        self._vb_cache = {}
        self._price_static = {}
        self.slab = None             # bind(shard=True): this rank's voxel rows (a slice); the plan holds only those
        self.group = None
        self.world = 1
        self._rmp_active = False     # solveRMC is running on the dense column cache
        self._gb_valid = False       # the plan's beamlet gradient belongs to the current dose

    # currentDose / voxelgradient are read back from the GPU only when somebody looks
    @property
    def currentDose(self):
        if self._dose is None and self.plan is not None and self._fg is not None:
            self._dose = self.plan.dose()
        return self._dose

    @currentDose.setter
    def currentDose(self, v):
        self._dose = v

    @property
    def voxelgradient(self):
        if self._vg is None and self.plan is not None and self._fg is not None:
            self._vg = self.plan.voxelgrad()
        return self._vg

    @voxelgradient.setter
    def voxelgradient(self, v):
        self._vg = v

    @property
    def beamletgradient(self):
        """D^T voxelgradient for every beamlet: each control point's ``beamGrad`` (:646)."""
        if self._gb is None and self.plan is not None and self._fg is not None:
            self._gb = self.plan.beamlet_grad()
        return self._gb

    def _sync_apertures(self):
        selected = getattr(self.caligraphicC, "loc", self.caligraphicC)      # the earlier generations keep a plain list
        sel = set(selected)
        for i in list(self._uploaded):
            if i not in sel:
                self.plan.set_aperture(i, None, None)
                del self._uploaded[i]
        for i in selected:
            # what was uploaded is remembered by object (holding the references, so an address cannot
            # be recycled for new arrays): colGen assigns fresh objects whenever an aperture changes
            key = (self.openApertureMaps[i], self.diagmakers[i], self.strengths[i])
            old = self._uploaded.get(i)
            if old is not None and all(a is b for a, b in zip(old, key)):
                continue
            nbl = int(self.beam_ptr[i + 1] - self.beam_ptr[i])
            w_dose = np.zeros(nbl)
            oam = np.asarray(self.openApertureMaps[i], dtype=np.int64)
            if oam.size:
                # a beamlet listed twice in the open map (:1089 restarts a row at index 0 when the
                # left leaf is left of the row's first beamlet) is deposited twice by calcDose (:250)
                np.add.at(w_dose, oam, np.asarray(self.strengths[i], dtype=float))
            w_grad = np.zeros(nbl)
            dm = np.asarray(self.diagmakers[i], dtype=float)
            w_grad[:len(dm)] = dm
            self.plan.set_aperture(i, w_dose, w_grad)
            self._uploaded[i] = key

    def calcDose(self):
        """Dose of the current apertures and intensities (greedyVMATsampling.py:243-251).
        On this path the same launch sequence also produces objective and gradients; they are
        published by calcGradientandObjValue()."""
        if self.plan is None:
            raise VmatError("no GPU plan bound: call VMATlibrary.bind(case) first")
        y = np.asarray(self.currentIntensities, dtype=float).ravel()
        if self._rmp_active:
            self._fg = self.plan.rmp_eval(y)         # apertures are fixed during an RMP: dense columns
            self._gb_valid = False
        else:
            self._sync_apertures()
            self._fg = self.plan.eval(y)
            self._gb_valid = True
        self._dose = self._vg = self._gb = None
        self.dZdK = _LazyDZdK(self)

    def calcGradientandObjValue(self):
        """Objective, voxel gradient, aperture gradient (greedyVMATsampling.py:254-272)."""
        if self._fg is None:
            self.calcDose()
        f, g = self._fg
        self.objectiveValue = np.float64(f)
        self.aperturegradient = np.asmatrix(g).transpose()


def bind(case, device: int = 0, store: str = "f32", acc: str = "f64", gpu: bool = True, group=None,
         shard: bool = False, **plan_opts):
    """Create the module-level ``data`` for a case (anything with the attributes of
    vmatwpencode_b200.synth.Case) and upload its D to GPU `device`.

    `acc="f64"` (default here) accumulates in double so that the greedy loop reproduces the
    reference's aperture sequence; `acc="f32"` is the throughput mode.

    `shard=True` (under torchrun, one process per GPU, torch.distributed initialised): this rank
    keeps only its voxel slab of D; every evaluation all-reduces the beamlet gradient and the
    objective inside the reduction kernel over peer memory, summing the ranks in rank order, so
    all ranks see bit-identical (f, g, beamlet gradient) and run the same greedy loop in lock
    step.  ``data.currentDose`` / ``voxelgradient`` are then the slab's."""
    global data, N, M
    d = vmat_class()
    d.numvoxels = case.V
    d.numbeams = case.nb
    d.beam_ptr = np.asarray(case.beam_ptr, dtype=np.int64)
    d.numX = int(d.beam_ptr[-1])
    d.beamletsPerBeam = np.diff(d.beam_ptr)
    d.xinter = np.asarray(case.xinter)
    d.yinter = np.asarray(case.yinter)
    d.xdirection = [np.asarray(a) for a in case.xdirection]
    d.ydirection = [np.asarray(a) for a in case.ydirection]
    d.pointtoAngle = list(case.angles)
    d.currentIntensities = np.zeros(case.nb, dtype=float)
    d.openApertureMaps = [[] for _ in range(case.nb)]
    d.diagmakers = [[] for _ in range(case.nb)]
    d.strengths = [[] for _ in range(case.nb)]
    d.llist = [[-1] * len(d.xinter) for _ in range(case.nb)]           # fully open (:952-954)
    d.rlist = [[len(d.yinter) + 1] * len(d.xinter) for _ in range(case.nb)]
    d.totaldijs = int(case.D.nnz)
    d.case_thr, d.case_over, d.case_under = np.asarray(case.thr), np.asarray(case.over), np.asarray(case.under)
    for name in ("numstructs", "numtargets", "numoars", "regions", "maskValue", "fullMaskValue", "regionIndices",
                 "targets", "oars", "functionData", "voxelAssignment"):       # present on ingested cases
        if hasattr(case, name):
            setattr(d, name, getattr(case, name))
    if gpu and shard:
        import torch
        import torch.distributed as dist
        from .dist import slab_of
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        rows, Dslab, thr, over, under = slab_of(case, rank, world)
        d.plan = DosePlan(Dslab, case.beam_ptr, thr, over, under, device=device, store=store, acc=acc, **plan_opts)
        d.slab = rows
        d.group = group
        d.world = world
        if world > 1:
            mine = torch.tensor(list(d.plan.comm_handle()), dtype=torch.uint8, device=f"cuda:{device}")
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine, group=group)
            d.plan.comm_attach(rank, world, b"".join(bytes(h.cpu().tolist()) for h in allh))
            dist.barrier(group)
    elif gpu:     # gpu=False: host-side bookkeeping only (geometry, updateOpenAperture); no evaluation
        d.plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, device=device,
                          store=store, acc=acc, **plan_opts)
    data = d
    N = len(d.yinter)
    M = len(d.xinter)
    return d


def quadHelpers(functionData, regions, numvoxels, priority=None):
    """Per-voxel quadHelperThresh / quadHelperOver / quadHelperUnder from the 3 x numstructs
    objective table (thresholds, over-weights, under-weights) and the structures' voxel lists
    (greedyVMATsampling.py:554-576): columns reordered by `priority`, weights divided by the
    structure's size, later structures overwrite earlier ones where voxel lists overlap."""
    fd = np.array(functionData, dtype=float).reshape(3, -1)
    if priority is not None:
        fd = fd[:, priority]
    thr = np.zeros(numvoxels)
    over = np.zeros(numvoxels)
    under = np.zeros(numvoxels)
    for s, reg in enumerate(regions):
        n = reg.sizeInVoxels
        if n > 0:
            idx = np.asarray(reg.indices, dtype=np.int64)
            thr[idx] = fd[0, s]
            over[idx] = fd[1, s] * 1 / n
            under[idx] = fd[2, s] * 1 / n
    return thr, over, under


def calcObjGrad(x, user_data=None):
    """(objective, aperture gradient) at intensities x (greedyVMATsampling.py:578-582)."""
    data.currentIntensities = x
    data.calcDose()
    data.calcGradientandObjValue()
    return data.objectiveValue, data.aperturegradient


def _beamlet_grid(index):
    """M x N occupancy of control point `index`: entry (i, y) says a beamlet sits at
    (xinter[i], yinter[y]) - the membership tests of :589-596 for all leaf rows at once."""
    hit = data._vb_cache.get(index)
    if hit is None:
        inrow = np.asarray(data.xdirection[index])[:, None] == np.asarray(data.xinter)[None, :]
        incol = np.asarray(data.ydirection[index])[:, None] == np.asarray(data.yinter)[None, :]
        hit = data._vb_cache[index] = (inrow.T.astype(np.int32) @ incol.astype(np.int32)) > 0
    return hit


def fvalidbeamlets(i, index):
    """Beamlet y-indices that exist in leaf row i of control point `index`, and the same
    range with both neighbours appended (greedyVMATsampling.py:589-602)."""
    vb = np.flatnonzero(_beamlet_grid(index)[i])
    if vb.size == 0:
        raise ValueError("min() arg is an empty sequence")       # what the reference raises
    return vb, np.concatenate(([vb[0] - 1], vb, [vb[-1] + 1]))


def _neighbour_leaves(predec, succ, Nn, Mm, vmax):
    """Leaf limits and leaf speeds of the adjacent selected apertures (:620-642)."""
    vmaxm = vmaxp = vmax
    if type(predec) is list:
        lcm, rcm, vmaxm = [-1] * Mm, [Nn] * Mm, float("inf")
    else:
        lcm, rcm = data.llist[predec], data.rlist[predec]
    if type(succ) is list:
        lcp, rcp, vmaxp = [-1] * Mm, [Nn] * Mm, float("inf")
    else:
        lcp, rcp = data.llist[succ], data.rlist[succ]
    return lcm, rcm, lcp, rcp, vmaxm, vmaxp


_PRICE_ROW_DTYPE = np.dtype([("l_first", "<f8"), ("n_l", "<i4"), ("vb_min", "<i4"), ("r_lo", "<f8"), ("r_hi", "<f8"),
                             ("r_mid", "<f8"), ("vb_mask", "<u8"), ("grad_offset", "<i4"), ("n_r_mid", "<i4")])


def _price_static(index, Mm):
    """Per control point, fixed by the beamlet grid: fvalidbeamlets of every leaf row as a bit mask,
    its minimum, and the running ``leftmostleaf`` offsets of :705,756 (quirk Q1: the first row
    counts one beamlet less)."""
    hit = data._price_static.get(index)
    if hit is None:
        grid = _beamlet_grid(index)[:Mm]
        count = grid.sum(axis=1)
        if (count == 0).any():
            raise ValueError("min() arg is an empty sequence")   # what the reference's fvalidbeamlets raises
        bits = np.uint64(1) << np.arange(grid.shape[1], dtype=np.uint64)
        mask = (grid * bits[None, :]).sum(axis=1, dtype=np.uint64)
        offs = np.zeros(Mm, dtype=np.int32)
        offs[1:] = np.cumsum(count)[:-1] - 1
        hit = data._price_static[index] = (mask, grid.argmax(axis=1).astype(np.int32), offs)
    return hit


def _fill_price_rows(rows, base, angdistancem, angdistancep, vmax, speedlim, predec, succ, Nn, Mm,
                     thisApertureIndex, bw):
    """Host part of PPsubroutine: the leaf windows of every row (:664,672-680,713-724) and the
    beamlet bookkeeping (fvalidbeamlets, the running ``leftmostleaf`` of :705,756).  Same float64
    expressions as the reference, evaluated for all M rows at once."""
    if Nn > 62:
        raise VmatError("pricing kernel supports at most 62 beamlets per leaf row")
    lcm, rcm, lcp, rcp, vmaxm, vmaxp = _neighbour_leaves(predec, succ, Nn, Mm, vmax)
    lcm, rcm, lcp, rcp = (np.asarray(v, dtype=np.float64) for v in (lcm, rcm, lcp, rcp))
    travelm = vmaxm * (angdistancem / speedlim) / bw
    travelp = vmaxp * (angdistancep / speedlim) / bw
    out = np.frombuffer(rows, dtype=_PRICE_ROW_DTYPE)[base:base + Mm]
    lo = np.ceil(np.maximum(-1.0, np.maximum(lcm - travelm, lcp - travelp)))
    hi = 1.0 + np.floor(np.minimum(Nn - 1.0, np.minimum(lcm + travelm, lcp + travelp)))
    empty = hi - lo <= 0
    out["l_first"] = lo
    out["n_l"] = (hi - lo).astype(np.int32)
    finite = not (math.isinf(angdistancem) or math.isinf(angdistancep))
    if finite and empty.any():
        # np.arange(mid, mid + 1) of the reference: one value, or two when mid + 1 rounds up
        mid = (angdistancep * lcm + angdistancem * lcp) / (angdistancep + angdistancem)
        out["l_first"][empty] = mid[empty]
        out["n_l"][empty] = np.ceil((mid[empty] + 1) - mid[empty]).astype(np.int32)
    out["r_lo"] = np.maximum(rcm - travelm, rcp - travelp)
    out["r_hi"] = np.minimum(float(Nn), np.minimum(rcm + travelm, rcp + travelp))
    if finite:
        rmid = (angdistancep * rcm + angdistancem * rcp) / (angdistancep + angdistancem)
        out["r_mid"] = rmid
        out["n_r_mid"] = np.ceil((rmid + 1) - rmid).astype(np.int32)
    else:
        out["r_mid"] = np.nan                 # unreachable: an infinite travel never empties a range
        out["n_r_mid"] = 1
    out["vb_mask"], out["vb_min"], out["grad_offset"] = _price_static(thisApertureIndex, Mm)


def _price(cands, C_, C2, C3, b, vmax, speedlim, Nn, Mm, bw):
    """Price a list of (index, predec, succ, angdistancem, angdistancep) in one launch."""
    if data._fg is None:
        raise VmatError("pricing needs the beamlet gradient of an evaluation: call data.calcDose() first")
    if not data._gb_valid:       # the last evaluations came from the RMP column cache: refresh D^T vg
        data.calcDose()
    rows = (PriceRow * (len(cands) * Mm))()
    for k, (idx, predec, succ, am, ap) in enumerate(cands):
        _fill_price_rows(rows, k * Mm, am, ap, vmax, speedlim, predec, succ, Nn, Mm, idx, bw)
    prm = PriceParams(C=float(C_), C2=float(C2), C3=float(C3), b=float(b), M=Mm, N=Nn)
    return data.plan.price([c[0] for c in cands], prm, rows)


def PPsubroutine(C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim, predec, succ, N, M,
                 thisApertureIndex, bw):
    """Best aperture for one control point: (p, l, r) with len(l) == len(r) == M
    (greedyVMATsampling.py:618-789).  Uses the beamlet gradient of the last evaluation."""
    p, l, r = _price([(thisApertureIndex, predec, succ, angdistancem, angdistancep)],
                     C, C2, C3, b, vmax, speedlim, N, M, bw)
    return np.float64(p[0]), [np.int64(v) for v in l[0]], [np.int64(v) for v in r[0]]


def _neighbours(thisApertureIndex):
    """Nearest selected predecessor / successor and angular gaps (:796-818)."""
    succs = [i for i in data.caligraphicC.loc if i > thisApertureIndex]
    predecs = [i for i in data.caligraphicC.loc if i < thisApertureIndex]
    if 0 == len(succs):
        succ, angdistancep = [], np.inf
    else:
        succ = min(succs)
        angdistancep = data.caligraphicC(succ) - data.notinC(thisApertureIndex)
    if 0 == len(predecs):
        predec, angdistancem = [], np.inf
    else:
        predec = max(predecs)
        angdistancem = data.notinC(thisApertureIndex) - data.caligraphicC(predec)
    return predec, succ, angdistancem, angdistancep


def parallelizationPricingProblem(i, C, C2, C3, b, vmax, speedlim, N, M, bw):
    """(p, l, r, i) for candidate control point i (greedyVMATsampling.py:791-822)."""
    predec, succ, am, ap = _neighbours(i)
    p, l, r = PPsubroutine(C, C2, C3, b, am, ap, vmax, speedlim, predec, succ, N, M, i, bw)
    return p, l, r, i


def PricingProblem(C, C2, C3, b, vmax, speedlim, N, M, bw):
    """Price every candidate in data.notinC and return the most negative one:
    (pstar, l, r, bestApertureIndex) (greedyVMATsampling.py:824-856).  The reference forks a
    process pool over candidates; here all candidates are one kernel launch."""
    cands = []
    for i in data.notinC.loc:
        predec, succ, am, ap = _neighbours(i)
        cands.append((i, predec, succ, am, ap))
    p, l, r = _price(cands, C, C2, C3, b, vmax, speedlim, N, M, bw)
    indstar = int(np.argmin(p))
    for k in range(len(cands)):
        if M != len(l[k]):
            sys.exit("Length of left vector is less than 28")
    if len(penalizationweights) == data.numbeams:         # Laptop variant bookkeeping (:863)
        penalizationweights[cands[indstar][0]] = kelly_measure(l[indstar], r[indstar])
    return (np.float64(p[indstar]), [np.int64(v) for v in l[indstar]], [np.int64(v) for v in r[indstar]],
            cands[indstar][0])


def kelly_measure(l, r):
    """Aperture complexity perimeter / area of leaf vectors (l, r): Kelly's measure as computed
    inside the Laptop variant's PricingProblem (greedyVMATsamplingLaptop.py:850-862)."""
    l = np.asarray(l, dtype=float)
    r = np.asarray(r, dtype=float)
    width = r - l
    area = 0.25 * (width.sum() + width[1:].sum())
    edge = np.sign(width)
    perimeter = width[0] + edge[0] + width[-1] + edge[-1] + edge[1:].sum()
    step = (np.abs(np.diff(l)) + np.abs(np.diff(r)) - 2 * np.maximum(0, l[:-1] - r[1:])
            - 2 * np.maximum(0, l[1:] - r[:-1]))
    return (perimeter + step.sum()) / area


def updateOpenAperture(i):
    """Leaf positions of control point i -> (open beamlet indices, diagmaker, strengths)
    (greedyVMATsampling.py:1070-1130), including its quirks: the first open beamlet of a row
    has strength ceil(left)-left (0 for integer leaves), and fractional right-edge strengths
    are written to the LAST element of the beam-wide diagmaker."""
    nbl = int(data.beam_ptr[i + 1] - data.beam_ptr[i])
    diagmaker = np.zeros(nbl, dtype=float)
    open_idx = []
    open_strength = []
    seen = 0
    for m in range(len(data.llist[i])):
        vb, _ = fvalidbeamlets(m, i)
        first_y, last_y, count = int(vb.min()), int(vb.max()), len(vb)
        lpos, rpos = data.llist[i][m], data.rlist[i][m]
        indleft = (lpos + 1 + seen - first_y) if lpos >= first_y - 1 else 0
        if rpos > last_y:
            indright = count + seen
        else:
            indright = (rpos - 1 + seen - first_y) if rpos >= first_y else 0
        seen += count
        k0, k1 = int(np.floor(indleft)), int(np.ceil(indright))
        if k0 >= k1:
            continue
        run = list(range(k0, k1))
        s = [1.0] * len(run)
        s[0] = np.ceil(indleft) - indleft
        for k, sv in zip(run, s):
            diagmaker[k] = sv
        open_idx.extend(run)
        open_strength.extend(s)
        edge = indright - np.floor(indright)
        if edge > 0.01:
            open_strength[-1] = edge
            diagmaker[-1] = edge
        if k1 - k0 == 1:
            open_strength[-1] = indright - indleft
            diagmaker[-1] = indright - indleft
    return np.array(open_idx, dtype=int), diagmaker, open_strength


def _lbfgsb_lean(fg, x0, lower, upper, ftol, maxiter, gtol=1e-5, maxcor=10, maxfun=15000, maxls=20):
    """The loop of scipy.optimize._lbfgsb_py._minimize_lbfgsb (scipy 1.18) around the same compiled routine
    (_lbfgsb.setulb) with the same work arrays and stopping rules, minus ScalarFunction, bounds objects and
    OptimizeResult bookkeeping per iteration.  `fg(x) -> (f, g)`; every variable has both bounds.  Returns an
    OptimizeResult with the fields the greedy loop reads (x, fun, jac, nit, nfev, status, success, message)."""
    from scipy.optimize import OptimizeResult, _lbfgsb
    n = len(x0)
    m = maxcor
    factr = ftol / np.finfo(float).eps
    idt = np.int64 if getattr(_lbfgsb, "HAS_ILP64", False) else np.int32
    try:
        from scipy.optimize._lbfgsb_py import HAS_ILP64
        idt = np.int64 if HAS_ILP64 else np.int32
    except ImportError:
        pass
    nbd = np.full(n, 2, dtype=idt)
    low = np.ascontiguousarray(lower, dtype=np.float64)
    up = np.ascontiguousarray(upper, dtype=np.float64)
    x = np.clip(np.array(x0, dtype=np.float64).ravel(), low, up)
    f = np.array(0.0, dtype=np.float64)
    g = np.zeros(n, dtype=np.float64)
    wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
    iwa = np.zeros(3 * n, dtype=idt)
    task = np.zeros(2, dtype=idt)
    ln_task = np.zeros(2, dtype=idt)
    lsave = np.zeros(4, dtype=idt)
    isave = np.zeros(44, dtype=idt)
    dsave = np.zeros(29, dtype=np.float64)
    nit = nfev = 0
    while True:
        _lbfgsb.setulb(m, x, low, up, nbd, f, g, factr, gtol, wa, iwa, task, lsave, isave, dsave, maxls, ln_task)
        if task[0] == 3:
            fv, gv = fg(x.copy())                      # (setulb moves x in place; the callback keeps what it was given)
            nfev += 1
            f = np.array(fv, dtype=np.float64)
            g = np.array(gv, dtype=np.float64)        # a fresh array, as minimize() hands over
        elif task[0] == 1:
            nit += 1
            if nit >= maxiter:
                task[0], task[1] = 5, 504
            elif nfev > maxfun:
                task[0], task[1] = 5, 502
        else:
            break
    warnflag = 0 if task[0] == 4 else (1 if (nfev > maxfun or nit >= maxiter) else 2)
    return OptimizeResult(fun=float(f), jac=g, nfev=nfev, njev=nfev, nit=nit, status=warnflag, x=x,
                          success=(warnflag == 0), message="lean L-BFGS-B loop over scipy.optimize._lbfgsb.setulb")


def _lean_available():
    try:
        from scipy.optimize import _lbfgsb
        import inspect
        import scipy
        return hasattr(_lbfgsb, "setulb") and tuple(int(v) for v in scipy.__version__.split(".")[:2]) >= (1, 15)
    except Exception:
        return False


def solveRMC(YU):
    """Restricted master problem: scipy L-BFGS-B over all control-point intensities, free
    only where an aperture is selected (greedyVMATsampling.py:858-885).  scipy stays on the
    host so that the iterates are the reference's; every callback is one GPU evaluation."""
    start = time.time()
    if fast_rmp and not use_column_cache and data.plan is not None and _lean_available():
        # The apertures are fixed during an RMP: upload them once, then every evaluation is the bare C call
        # (host buffers in and out), handed to scipy's compiled L-BFGS-B routine by a lean loop.
        plan = data.plan
        data._sync_apertures()
        lib_eval, args, ybuf, gbuf, fcell = plan._lib.vmat_eval, plan._eval_args, plan._y, plan._g, plan._f

        def fg(x):
            data.currentIntensities = x
            ybuf[:] = x
            rc = lib_eval(*args)
            if rc:
                from ._lib import check
                check(rc, "vmat_eval")
            return fcell.value, gbuf

        sel = set(data.caligraphicC.loc)
        upper = np.array([YU if k in sel else 0.0 for k in range(data.numbeams)])
        f0, g0 = fg(np.asarray(data.currentIntensities, dtype=float).ravel())        # the reference's warm-up evaluation (:873)
        res = _lbfgsb_lean(fg, data.currentIntensities, np.zeros(data.numbeams), upper, ftol=1e-1, maxiter=200)
        # what the last callback leaves behind in the reference: objective and gradient of the last point evaluated
        data._fg = (fcell.value, gbuf.copy())
        data._gb_valid = True
        data._dose = data._vg = data._gb = None
        data.dZdK = _LazyDZdK(data)
        data.calcGradientandObjValue()
        data.rmp_seconds = time.time() - start
        return res
    cached = bool(use_column_cache and data.plan is not None and data.caligraphicC.len() > 0
                  and getattr(data.plan, "fused_exchange", False) is False)
    if cached:
        # The apertures do not change inside an RMP, so D_i^T strengths_i and D_i^T diagmaker_i are
        # computed once (sparse kernel) and every L-BFGS-B callback works on those dense columns.
        data._sync_apertures()
        data.plan.rmp_begin(list(data.caligraphicC.loc))
        data._rmp_active = True
    try:
        calcObjGrad(data.currentIntensities)
        boundschoice = [(0, YU) if k in data.caligraphicC.loc else (0, 0) for k in range(data.numbeams)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = minimize(calcObjGrad, data.currentIntensities, method="L-BFGS-B", jac=True,
                           bounds=boundschoice, options={"ftol": 1e-1, "maxiter": 200})
    finally:
        if cached:
            data._rmp_active = False
            data.plan.rmp_end()
    data.rmp_seconds = time.time() - start
    return res


def printresults(iterationNumber, myfolder=None, Cvalue=None):
    """Dose-volume histograms of the current dose: (bin_center, dvh_matrix[numstructs, bins]) - the
    numbers the reference's printresults plots (greedyVMATsampling.py:888-912; Laptop :902-926).
    Dose maximum and histograms are computed on the GPU from the resident dose; nothing is
    plotted.  With `myfolder` the two arrays are saved there as an .npz."""
    if data.plan is None or data._fg is None:
        raise VmatError("printresults needs an evaluated plan: call data.calcDose() first")
    maxDose = data.plan.dose_max()
    sharded = data.slab is not None and data.world > 1
    if sharded:          # every rank holds a voxel slab: the maximum and the counts are combined over the ranks
        import torch
        import torch.distributed as dist
        t = torch.tensor([maxDose], dtype=torch.float64, device=f"cuda:{data.plan.device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=data.group)
        maxDose = float(t.item())
    dose_resln = 0.1
    dose_ub = maxDose + 10
    bin_center = np.arange(0, dose_ub, dose_resln)
    full = np.array([int(v) for v in data.fullMaskValue], dtype=np.uint64)
    hist = data.plan.dvh_histogram(full[data.slab] if data.slab is not None else full, data.numstructs, bin_center)
    if sharded:
        t = torch.as_tensor(hist, device=f"cuda:{data.plan.device}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=data.group)
        hist = t.cpu().numpy()
    dvh_matrix = np.zeros((data.numstructs, len(bin_center)))
    for s in range(data.numstructs):
        if hist[s].sum() == 0 and not np.any(full & np.uint64(2 ** s)):
            continue                                              # empty structure: row stays 0 (:904-905)
        cum = np.cumsum(np.append(hist[s], 0))
        dvh_matrix[s, :] = 1 - cum / max(cum)
    if myfolder:
        tag = "" if Cvalue is None else "-C" + str(Cvalue)
        np.savez(os.path.join(myfolder, "DVH-iteration-" + str(iterationNumber) + tag + ".npz"),
                 bin_center=bin_center, dvh_matrix=dvh_matrix)
    return bin_center, dvh_matrix


RESULT_FIELDS = ("numbeams", "x", "C", "C2", "C3", "vmax", "speedlim", "RU", "YU", "M", "N", "llist", "rlist",
                 "fullMaskValue", "currentDose", "currentIntensities", "numstructs", "allNames", "objectiveValue",
                 "quadHelperThresh", "quadHelperOver", "quadHelperUnder", "regionIndices", "targets", "oars")


def saveResults(PIK, C, C2=1.0, C3=1.0, vmax=2.25, speedlim=0.83, RU=10.0, allNames=None):
    """Write the 25-field result list colGen pickles at the end of a run
    (greedyVMATsampling.py:1053-1060) in the order picklereader.py:108-134 unpacks it."""
    import pickle
    plan = data.plan
    thr = over = under = None
    if plan is not None:
        thr, over, under = data.case_thr, data.case_over, data.case_under
    dose = data.currentDose
    if data.slab is not None and data.world > 1:      # sharded: the bundle holds the whole dose, slabs in rank order
        import torch.distributed as dist
        parts = [None] * data.world
        dist.all_gather_object(parts, np.asarray(dose), group=data.group)
        dose = np.concatenate(parts)
    datasave = [data.numbeams, data.rmpres.x, C, C2, C3, vmax, speedlim, RU, RU / speedlim, M, N, data.llist, data.rlist,
                data.fullMaskValue, dose, data.currentIntensities, data.numstructs,
                list(allNames) if allNames is not None else [], data.objectiveValue, thr, over, under,
                data.regionIndices, data.targets, data.oars]
    with open(PIK, "wb") as f:
        pickle.dump(datasave, f, pickle.HIGHEST_PROTOCOL)
    return datasave


def loadResults(PIK):
    """Read a result bundle back as a dict keyed by the names picklereader.py uses."""
    import pickle
    with open(PIK, "rb") as f:
        items = pickle.load(f)
    return dict(zip(RESULT_FIELDS, items))


def colGen(C, WholeCircle, initialApertures, max_iter=None):
    """Greedy column generation (greedyVMATsampling.py:937-1061): price, add the best
    aperture, re-optimise intensities, resample candidates; returns pstar.  Each step is
    appended to data.history.  (`max_iter` is an addition for bounded test runs.)"""
    global kappa, penalizationweights
    C2 = 1.0
    C3 = 1.0
    penalizationweights = [None] * data.numbeams
    vmax = 2.25
    speedlim = 0.83
    beamletwidth = 0.5
    RU = 10.0
    YU = RU / speedlim
    data.llist = [[-1] * len(data.xinter) for _ in range(data.numbeams)]
    data.rlist = [[len(data.yinter) + 1] * len(data.xinter) for _ in range(data.numbeams)]
    if not WholeCircle or not kappa:
        kappa = list(range(data.numbeams))
    random.seed(13)
    kappa = random.sample(kappa, len(kappa))
    for _ in range(min(initialApertures, len(kappa))):
        i = kappa.pop(0)
        data.notinC.insertAngle(i, data.pointtoAngle[i])
    data.caligraphicC = apertureList()
    pstar = -np.inf
    data.history = []
    while (pstar < 0) & (data.notinC.len() > 0):
        data.calcDose()
        data.calcGradientandObjValue()
        pstar, lm, rm, bestApertureIndex = PricingProblem(C, C2, C3, 0.5, vmax, speedlim, N, M, beamletwidth)
        if pstar >= 0:
            break
        data.caligraphicC.insertAngle(bestApertureIndex, data.notinC(bestApertureIndex))
        data.notinC.removeIndex(bestApertureIndex)
        data.llist[bestApertureIndex] = lm
        data.rlist[bestApertureIndex] = rm
        (data.openApertureMaps[bestApertureIndex], data.diagmakers[bestApertureIndex],
         data.strengths[bestApertureIndex]) = updateOpenAperture(bestApertureIndex)
        rmpres = solveRMC(YU)
        data.rmpres = rmpres
        removed = []
        for thisindex in range(data.numbeams):
            if thisindex in data.caligraphicC.loc:
                if (rmpres.x[thisindex] < eliminationThreshold) & eliminationPhase:
                    data.entryCounter += 1
                    removed.append(thisindex)
                    data.notinC.insertAngle(thisindex, data.pointtoAngle[thisindex])
                    data.caligraphicC.removeIndex(thisindex)
                    penalizationweights[thisindex] = None
        data.history.append(dict(idx=int(bestApertureIndex), l=[int(v) for v in lm], r=[int(v) for v in rm],
                                 pstar=float(pstar), x=np.array(rmpres.x, dtype=float), fun=float(rmpres.fun),
                                 nit=int(rmpres.nit), nfev=int(rmpres.nfev), removed=list(removed),
                                 rmp_s=float(getattr(data, "rmp_seconds", 0.0)),
                                 kelly=float(kelly_measure(lm, rm))))
        if len(data.listIndexofAperturesRemovedEachStep) > 1:
            # the same aperture keeps entering and leaving: stop (:1020-1031)
            if np.any(np.isin(removed, data.listIndexofAperturesRemovedEachStep[-1])):
                break
            if np.any(np.isin(removed, data.listIndexofAperturesRemovedEachStep[-2])):
                break
        data.listIndexofAperturesRemovedEachStep.append(removed)
        while not data.notinC.isEmpty():
            kappa.append(data.notinC.loc[0])
            data.notinC.removeIndex(data.notinC.loc[0])
        random.seed(13)
        for i in random.sample(kappa, min(initialApertures, len(kappa))):
            data.notinC.insertAngle(i, data.pointtoAngle[i])
            kappa.remove(i)
        if max_iter is not None and len(data.history) >= max_iter:
            break
    return pstar
