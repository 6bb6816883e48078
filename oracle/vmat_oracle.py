"""TEST INFRASTRUCTURE - CPU restatement (numpy/scipy, float64) of the reference hot path.

This is the oracle the CUDA path is checked against.  It is NOT the product:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import it.  The product (vmatwpencode_b200) never does and fails
loudly when its CUDA library is missing.

Parity pinning: the reference has no tests or golden vectors for the
dose/gradient/pricing path.  This restatement is
pinned by running the reference's own functions, AST-extracted from
/root/reference (oracle/ref_extract.py), on seeded synthetic cases in the
build container: tests/test_oracle_vs_reference.py compares live, and
oracle/make_golden.py froze the reference's outputs into tests/golden/*.npz
so the pin travels to machines without /root/reference.  For k-medoids the
reference embeds one known answer (the 178-entry ``kappa`` list,
greedyVMAT.py:28), which tests/test_kmedoids.py checks.

Every function cites the reference lines it follows.  Variant restated:
greedyVMATsampling.py (the canonical, most evolved variant).
"""
from __future__ import annotations

import math
import random
import warnings
from typing import List, Optional

import numpy as np
from scipy import sparse
from scipy.optimize import minimize


# --------------------------------------------------------------------------
# state containers (greedyVMATsampling.py:97-140 apertureList, :145-276 vmat_class)
# --------------------------------------------------------------------------
class ApertureSet:
    """Sorted (loc, angle) pairs; restates apertureList (greedyVMATsampling.py:97-140).

    Like the reference, loc and angle are sorted independently on insert, which
    keeps them paired only because angle is monotone in loc."""

    def __init__(self):
        self.loc: List[int] = []
        self.angle: List[float] = []

    def insertAngle(self, i, aperangle):
        self.loc.append(i)
        self.angle.append(aperangle)
        self.loc.sort()
        self.angle.sort()

    def removeIndex(self, index):
        k = self.loc.index(index)
        self.loc.pop(k)
        self.angle.pop(k)

    def __call__(self, index):
        return self.angle[self.loc.index(index)]

    def len(self):
        return len(self.loc)

    def isEmpty(self):
        return len(self.loc) == 0


class OracleState:
    """The fields of the reference's singleton ``data`` that the hot path touches."""

    def __init__(self, case):
        self.case = case
        self.numvoxels = case.V
        self.numbeams = case.nb
        self.M, self.N = case.M, case.N
        self.xinter = np.asarray(case.xinter)
        self.yinter = np.asarray(case.yinter)
        self.xdirection = [np.asarray(a) for a in case.xdirection]
        self.ydirection = [np.asarray(a) for a in case.ydirection]
        self.Dlist = case.Dlist()                      # B_i x V CSR  (:482,524)
        self.DlistT = [d.transpose() for d in self.Dlist]   # V x B_i CSC view (:525)
        self.thr = np.asarray(case.thr, dtype=float)
        self.over = np.asarray(case.over, dtype=float)
        self.under = np.asarray(case.under, dtype=float)
        self.pointtoAngle = list(case.angles)
        self.llist = [[-1] * case.M for _ in range(case.nb)]          # :952-954
        self.rlist = [[case.N + 1] * case.M for _ in range(case.nb)]
        self.openApertureMaps = [[] for _ in range(case.nb)]
        self.diagmakers = [[] for _ in range(case.nb)]
        self.strengths = [[] for _ in range(case.nb)]
        self.currentIntensities = np.zeros(case.nb)
        self.currentDose = np.zeros(case.V)
        self.voxelgradient = np.zeros(case.V)
        self.aperturegradient = None
        self.objectiveValue = float("inf")
        self.notinC = ApertureSet()
        self.caligraphicC = ApertureSet()
        self.removed_each_step = []                    # listIndexofAperturesRemovedEachStep
        # global ("clean") form: one V x B matrix and two beamlet weight vectors
        self.D = case.D
        self.beam_ptr = np.asarray(case.beam_ptr, dtype=np.int64)

    # -- global weight vectors w_dose / w_grad ----------------
    def weight_vectors(self):
        B = int(self.beam_ptr[-1])
        w_dose = np.zeros(B)
        w_grad = np.zeros(B)
        for i in self.caligraphicC.loc:
            lo = int(self.beam_ptr[i])
            oam = np.asarray(self.openApertureMaps[i], dtype=np.int64)
            if oam.size:
                # the open map can list a beamlet more than once (updateOpenAperture restarts a row
                # at index 0 when its left leaf lies left of the row's first beamlet, :1089); the
                # reference's column fancy-indexing then deposits that beamlet once per occurrence
                np.add.at(w_dose, lo + oam, np.asarray(self.strengths[i], dtype=float))
            w_grad[lo:lo + len(self.diagmakers[i])] = self.diagmakers[i]
        return w_dose, w_grad


def quad_helpers(functionData, regions, numvoxels, priority=None):
    """quadHelperThresh/Over/Under as the reference builds them at load time
    (greedyVMATsampling.py:554-576): per structure s (after the priority reordering) the
    weights are divided by the structure size, then written voxel by voxel."""
    fd = np.array([float(v) for v in np.ravel(functionData)]).reshape(3, -1)
    if priority is not None:
        fd = fd[:, priority]
    for s in range(fd.shape[1]):
        if regions[s].sizeInVoxels > 0:
            fd[1, s] = fd[1, s] * 1 / regions[s].sizeInVoxels
            fd[2, s] = fd[2, s] * 1 / regions[s].sizeInVoxels
    thr, over, under = np.zeros(numvoxels), np.zeros(numvoxels), np.zeros(numvoxels)
    for s in range(fd.shape[1]):
        for j in range(regions[s].sizeInVoxels):
            thr[int(regions[s].indices[j])] = fd[0][s]
            over[int(regions[s].indices[j])] = fd[1][s]
            under[int(regions[s].indices[j])] = fd[2][s]
    return thr, over, under


# --------------------------------------------------------------------------
# evaluation: dose, objective, gradients
# --------------------------------------------------------------------------
def calc_dose_literal(st: OracleState):
    """calcDose as the reference performs it (greedyVMATsampling.py:243-251): per selected
    control point, CSC column slice x diag(strengths) x repeat(y_i); dZdK column = row sums
    of D_i^T diag(diagmaker).  Returns (dose, dZdK dense V x nb)."""
    dose = np.zeros(st.numvoxels, dtype=float)
    dZdK = np.zeros((st.numvoxels, st.numbeams))
    for i in st.caligraphicC.loc:
        oam = st.openApertureMaps[i]
        if len(oam):
            scaled = st.DlistT[i][:, oam] @ sparse.diags(np.asarray(st.strengths[i], dtype=float))
            dose += scaled @ np.repeat(st.currentIntensities[i], len(oam))
        col = (st.DlistT[i] @ sparse.diags(np.asarray(st.diagmakers[i], dtype=float), 0)).sum(axis=1)
        dZdK[:, i] = np.asarray(col).ravel()
    return dose, dZdK


def penalty(dose, thr, over, under):
    """calcGradientandObjValue without the aperture part (greedyVMATsampling.py:254-271).
    Objective by Python's left-to-right ``sum`` (quirk Q8); voxel gradient carries the
    reference's factor 4 (quirk Q4).  Returns (objective, voxelgradient)."""
    o = dose - thr
    o = (o > 0) * o
    u = thr - dose
    u = (u > 0) * u
    obj = sum(o * o * over + u * u * under)
    vg = 2 * (2 * o * over - 2 * u * under)
    return obj, vg


def calc_obj_grad_literal(st: OracleState, x):
    """calcObjGrad (greedyVMATsampling.py:578-582) in the reference's own operation order.
    Returns (f, g) with g of shape (nb, 1) like the reference's np.matrix."""
    st.currentIntensities = x
    st.currentDose, dZdK = calc_dose_literal(st)
    st.objectiveValue, st.voxelgradient = penalty(st.currentDose, st.thr, st.over, st.under)
    st.aperturegradient = (st.voxelgradient[None, :] @ dZdK).T
    return st.objectiveValue, st.aperturegradient


def calc_obj_grad_clean(st: OracleState, x, w_dose=None, w_grad=None):
    """Same evaluation in the global form the kernels implement:
    x_j = y[beam(j)] w_dose[j]; d = D x; f, vg; gb = D^T vg; g_i = sum_j w_grad[j] gb[j].
    Returns (f, g[nb], d, vg, gb)."""
    if w_dose is None:
        w_dose, w_grad = st.weight_vectors()
    st.currentIntensities = x
    nbeamlets = np.diff(st.beam_ptr)
    xb = np.repeat(np.asarray(x, dtype=float), nbeamlets) * w_dose
    d = st.D @ xb
    f, vg = penalty(d, st.thr, st.over, st.under)
    gb = st.D.T @ vg
    g = np.add.reduceat(w_grad * gb, st.beam_ptr[:-1]) if len(gb) else np.zeros(st.numbeams)
    g = np.where(nbeamlets > 0, g, 0.0)
    st.currentDose, st.voxelgradient, st.objectiveValue = d, vg, f
    st.aperturegradient = g.reshape(-1, 1)
    return f, g, d, vg, gb


# --------------------------------------------------------------------------
# beamlet geometry
# --------------------------------------------------------------------------
def valid_beamlets(st: OracleState, row: int, beam: int):
    """y indices of the beamlets that exist in leaf row `row` of control point `beam`
    (fvalidbeamlets, greedyVMATsampling.py:589-602)."""
    here = st.xdirection[beam] == st.xinter[row]
    present = np.isin(st.yinter, st.ydirection[beam][here])
    return np.flatnonzero(present)


def leaf_window(N, M, lcm, lcp, rcm, rcp, travelm, travelp, angm, angp):
    """Per-row leaf ranges of the pricing network (greedyVMATsampling.py:664,672-680,713-724).

    For each row returns a list of (l, [r...]) in enumeration order.  l and r are ints
    from ``range`` normally, or a single float midpoint when the speed constraints from
    the neighbouring apertures cannot both be met (quirk Q7)."""
    rows = []
    for m in range(M):
        lo = math.ceil(max(-1, lcm[m] - travelm, lcp[m] - travelp))
        hi = 1 + math.floor(min(N - 1, lcm[m] + travelm, lcp[m] + travelp))
        ls = list(range(lo, hi))
        if len(ls) == 0:
            mid = (angp * lcm[m] + angm * lcp[m]) / (angp + angm)
            ls = list(np.arange(mid, mid + 1))
        nodes = []
        for l in ls:
            rlo = math.ceil(max(l + 1, rcm[m] - travelm, rcp[m] - travelp))
            rhi = 1 + math.floor(min(N, rcm[m] + travelm, rcp[m] + travelp))
            rs = list(range(rlo, rhi))
            if len(rs) == 0:
                mid = (angp * rcm[m] + angm * rcp[m]) / (angp + angm)
                rs = list(np.arange(mid, mid + 1))
            nodes.append((l, rs))
        rows.append(nodes)
    return rows


# --------------------------------------------------------------------------
# pricing: layered shortest path over (row, l, r) nodes
# --------------------------------------------------------------------------
def pp_subroutine(st: OracleState, C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim,
                  predec, succ, N, M, thisApertureIndex, bw, beamGrad=None):
    """PPsubroutine (greedyVMATsampling.py:618-789).  Returns (p, l[M], r[M]).

    Restated row-at-a-time: all nodes of a row are enumerated first, their
    prev-independent terms computed, then the K_prev x K cost matrix is formed with the
    same elementwise float64 operation order as the reference's per-node vector code, so
    results are bitwise those of the reference."""
    vmaxm = vmaxp = vmax
    if isinstance(predec, list):                      # :625-632 no predecessor
        lcm, rcm, vmaxm = [-1] * M, [N] * M, float("inf")
    else:
        lcm, rcm = st.llist[predec], st.rlist[predec]
    if isinstance(succ, list):                        # :635-642
        lcp, rcp, vmaxp = [-1] * M, [N] * M, float("inf")
    else:
        lcp, rcp = st.llist[succ], st.rlist[succ]
    if beamGrad is None:
        beamGrad = st.Dlist[thisApertureIndex] @ st.voxelgradient      # :646
    travelm = vmaxm * (angdistancem / speedlim) / bw
    travelp = vmaxp * (angdistancep / speedlim) / bw
    window = leaf_window(N, M, lcm, lcp, rcm, rcp, travelm, travelp, angdistancem, angdistancep)

    prev_l = prev_r = prev_w = None
    dads = []          # per row >= 1: argmin index into the previous row
    row_l, row_r = [], []
    offset = 0         # "leftmostleaf" bookkeeping (:705,756): quirk Q1
    for m in range(M):
        vb = valid_beamlets(st, m, thisApertureIndex)
        vmin = int(vb.min())
        ls, rs, selfw, dose_t, c3s, tie, side = [], [], [], [], [], [], []
        for l, rlist in window[m]:
            for r in rlist:
                lo_i, hi_i = int(np.ceil(l + 1)), int(np.floor(r))
                inside = np.intersect1d(range(lo_i, hi_i), vb)
                ds = -((np.ceil(l + 1) - (l + 1)) * beamGrad[int(np.floor(l + 1))]
                       + (r - np.floor(r)) * beamGrad[int(np.ceil(r))])           # :690,736 (Q5)
                if m == 0:
                    if l + 1 < r:                                                  # :686-697
                        idx = inside - vmin
                        if len(idx) > 0:
                            dose = -beamGrad[idx].sum()
                            wgt = C * (C2 * (r - l) - C3 * b * (r - l)) - dose + 10E-10 * (r - l) + ds
                        else:
                            wgt = C * (C2 * (r - l) - C3 * b * (r - l)) + 10E-10 * (r - l) + ds
                    else:
                        wgt = 0.0
                    selfw.append(wgt)
                else:
                    idx = inside + offset - vmin                                   # :735 (Q1)
                    if len(idx) > 0:
                        dose_t.append(-beamGrad[idx].sum())
                        c3s.append(C3 * b * (r - l))
                    else:
                        dose_t.append(0.0)
                        c3s.append(0.0)
                    tie.append(10E-10 * (r - l))
                    side.append(ds)
                ls.append(l)
                rs.append(r)
        # what the int lnetwork/rnetwork arrays keep (:657-658,700-701,730-731): truncation (Q7)
        li = np.array([int(v) for v in ls], dtype=np.int64)
        ri = np.array([int(v) for v in rs], dtype=np.int64)
        if m == 0:
            w = np.array(selfw, dtype=float)
            offset = len(vb) - 1                                                   # :705
        else:
            lf = np.array(ls, dtype=float)[None, :]
            rf = np.array(rs, dtype=float)[None, :]
            pl = prev_l[:, None]
            pr = prev_r[:, None]
            lam = (np.absolute(pl - lf) + np.absolute(pr - rf)
                   - 2 * np.maximum(0, pl - rf) - 2 * np.maximum(0, lf - np.absolute(pr)))   # :743
            wgt = (C * (C2 * lam - np.array(c3s)[None, :]) - np.array(dose_t)[None, :]
                   + np.array(tie)[None, :] + np.array(side)[None, :])             # :744
            tot = prev_w[:, None] + wgt                                            # :746
            arg = np.argmin(tot, axis=0)                                           # :748 first minimum
            w = tot[arg, np.arange(tot.shape[1])]
            dads.append(arg)
            offset = len(vb) + offset                                              # :756
        prev_l, prev_r, prev_w = li, ri, w
        row_l.append(li)
        row_r.append(ri)

    # sink (:763-770): ``<=`` keeps the LAST minimum (quirk Q6)
    best, p = None, np.inf
    for k in range(len(prev_w)):
        cand = prev_w[k] + C * (C2 * (prev_r[k] - prev_l[k]))
        if cand <= p:
            p, best = cand, k
    # backtrack (:771-786)
    l_out, r_out = [0] * M, [0] * M
    k = best
    for m in range(M - 1, -1, -1):
        l_out[m] = int(row_l[m][k])
        r_out[m] = int(row_r[m][k])
        if m > 0:
            k = int(dads[m - 1][k])
    return p, l_out, r_out


def neighbours(st: OracleState, idx):
    """Nearest selected predecessor / successor and the angular gaps
    (parallelizationPricingProblem, greedyVMATsampling.py:791-818)."""
    succs = [i for i in st.caligraphicC.loc if i > idx]
    predecs = [i for i in st.caligraphicC.loc if i < idx]
    if succs:
        succ = min(succs)
        angp = st.caligraphicC(succ) - st.notinC(idx)
    else:
        succ, angp = [], np.inf
    if predecs:
        predec = max(predecs)
        angm = st.notinC(idx) - st.caligraphicC(predec)
    else:
        predec, angm = [], np.inf
    return predec, succ, angm, angp


def price_candidate(st, i, C, C2, C3, b, vmax, speedlim, N, M, bw):
    """parallelizationPricingProblem (greedyVMATsampling.py:791-822) -> (p, l, r, i)."""
    predec, succ, angm, angp = neighbours(st, i)
    p, l, r = pp_subroutine(st, C, C2, C3, b, angm, angp, vmax, speedlim, predec, succ, N, M, i, bw)
    return p, l, r, i


def pricing_problem(st, C, C2, C3, b, vmax, speedlim, N, M, bw):
    """PricingProblem (greedyVMATsampling.py:824-856), candidates priced serially in
    notinC.loc order, np.argmin (first minimum) picks the winner."""
    res = [price_candidate(st, i, C, C2, C3, b, vmax, speedlim, N, M, bw) for i in st.notinC.loc]
    k = int(np.argmin(np.array([r[0] for r in res])))
    for r in res:
        if len(r[1]) != M:
            raise SystemExit("Length of left vector is less than M")
    return res[k][0], res[k][1], res[k][2], res[k][3]


# --------------------------------------------------------------------------
# aperture -> beamlet weights
# --------------------------------------------------------------------------
def update_open_aperture(st: OracleState, i):
    """updateOpenAperture (greedyVMATsampling.py:1070-1130): leaf positions of control
    point i -> (open beamlet indices, diagmaker[B_i], strengths), quirks Q2/Q3 included."""
    nbl = st.Dlist[i].shape[0]
    diag = np.zeros(nbl, dtype=float)
    opened, strength_list = [], []
    base = 0
    for m in range(len(st.llist[i])):
        vb = valid_beamlets(st, m, i)
        vmin, vmax_ = int(vb.min()), int(vb.max())
        lm, rm = st.llist[i][m], st.rlist[i][m]
        left = lm + 1 + base - vmin if lm >= vmin - 1 else 0                       # :1083-1089
        if rm > vmax_:                                                             # :1091-1100
            right = len(vb) + base
        elif rm >= vmin:
            right = rm - 1 + base - vmin
        else:
            right = 0
        base += len(vb)
        if np.floor(left) < np.ceil(right):
            first = True
            for k in range(int(np.floor(left)), int(np.ceil(right))):
                s = 1.0
                if first:
                    first = False
                    s = np.ceil(left) - left                                       # :1112 (Q2)
                strength_list.append(s)
                diag[k] = s
                opened.append(k)
            s = right - np.floor(right)
            if s > 0.01:
                strength_list[-1] = s
                diag[-1] = s                                                       # :1122 (Q3)
            if 1 == int(np.ceil(right)) - int(np.floor(left)):
                s = right - left
                strength_list[-1] = s
                diag[-1] = s                                                       # :1128 (Q3)
    return np.array(opened, dtype=int), diag, strength_list


# --------------------------------------------------------------------------
# restricted master problem and the greedy loop
# --------------------------------------------------------------------------
def solve_rmc(st: OracleState, YU, evaluator=None):
    """solveRMC (greedyVMATsampling.py:858-885): scipy L-BFGS-B over all nb intensities,
    bounds (0, YU) for selected control points and (0, 0) otherwise."""
    fun = evaluator or (lambda x: calc_obj_grad_literal(st, x))
    fun(st.currentIntensities)
    bounds = [(0, YU) if k in st.caligraphicC.loc else (0, 0) for k in range(st.numbeams)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return minimize(fun, st.currentIntensities, method="L-BFGS-B", jac=True, bounds=bounds,
                        options={"ftol": 1e-1, "maxiter": 200})


def dvh_matrix(dose, full_mask, numstructs):
    """The numbers printresults plots (greedyVMATsampling.py:888-912): per structure s (bit s of
    the full mask) the cumulative dose-volume histogram on 0.1-wide bins up to max dose + 10."""
    mask = np.array([int(i) for i in full_mask])
    max_dose = max([float(i) for i in dose])
    bin_center = np.arange(0, max_dose + 10, 0.1)
    out = np.zeros((numstructs, len(bin_center)))
    for s in range(numstructs):
        holder = sorted(np.asarray(dose)[[i for i, v in enumerate(mask & 2 ** s) if v > 0]])
        if 0 == len(holder):
            continue
        hist, _ = np.histogram(holder, bin_center)
        hist = np.cumsum(np.append(hist, 0))
        out[s, ] = 1 - (hist / max(hist))
    return bin_center, out


def kelly_measure(l, r):
    """Aperture complexity (perimeter / area) of leaf vectors l, r
    (greedyVMATsamplingLaptop.py:850-862, inside its PricingProblem)."""
    area = 0.0
    perimeter = r[0] - l[0] + np.sign(r[0] - l[0])
    for n in range(len(l)):
        area += 0.5 * (r[n] - l[n]) * 0.5
    for n in range(1, len(l)):
        area += 0.5 * (r[n] - l[n]) * 0.5
        perimeter += np.sign(r[n] - l[n])
        perimeter += (np.abs(l[n] - l[n - 1]) + np.abs(r[n] - r[n - 1]) - 2 * np.maximum(0, l[n - 1] - r[n])
                      - 2 * np.maximum(0, l[n] - r[n - 1]))
    perimeter += r[len(r) - 1] - l[len(l) - 1] + np.sign(r[len(r) - 1] - l[len(l) - 1])
    return perimeter / area


def col_gen(st: OracleState, C, kappasize, max_iter=None, C2=1.0, C3=1.0, b=0.5, vmax=2.25,
            speedlim=0.83, bw=0.5, RU=10.0, clean=False, elimination=False, threshold=10E-3, kappa0=None):
    """colGen (greedyVMATsampling.py:937-1046) without plots/pickle.  `kappa0`: the module-level candidate
    ordering of the WholeCircle branch (:959-966; the reference hard-codes its 178-entry k-medoid ordering, :30);
    default: the non-WholeCircle branch (0 .. numbeams-1, :968-980).
    Returns the greedy history: per iteration idx, l, r, pstar, x, fun."""
    YU = RU / speedlim
    kappa = list(range(st.numbeams)) if kappa0 is None else list(kappa0)
    random.seed(13)
    kappa = random.sample(kappa, len(kappa))
    for _ in range(min(kappasize, len(kappa))):
        i = kappa.pop(0)
        st.notinC.insertAngle(i, st.pointtoAngle[i])
    st.caligraphicC = ApertureSet()
    if clean:
        def evaluator(x):
            f, g, *_ = calc_obj_grad_clean(st, x)
            return f, g
    else:
        evaluator = None
    pstar = -np.inf
    history = []
    while (pstar < 0) and (st.notinC.len() > 0):
        (evaluator or (lambda x: calc_obj_grad_literal(st, x)))(st.currentIntensities)   # :987-988
        pstar, lm, rm, best = pricing_problem(st, C, C2, C3, b, vmax, speedlim, st.N, st.M, bw)
        if pstar >= 0:
            break
        st.caligraphicC.insertAngle(best, st.notinC(best))
        st.notinC.removeIndex(best)
        st.llist[best], st.rlist[best] = lm, rm
        st.openApertureMaps[best], st.diagmakers[best], st.strengths[best] = update_open_aperture(st, best)
        res = solve_rmc(st, YU, evaluator)
        removed = []                                                               # :1007-1033
        for k in range(st.numbeams):
            if k in st.caligraphicC.loc and (res.x[k] < threshold) and elimination:
                removed.append(k)
                st.notinC.insertAngle(k, st.pointtoAngle[k])
                st.caligraphicC.removeIndex(k)
        prev = st.removed_each_step
        stop = len(prev) > 1 and bool(np.any(np.isin(removed, prev[-1])) or np.any(np.isin(removed, prev[-2])))
        if not stop:
            prev.append(removed)
        history.append(dict(idx=int(best), l=list(lm), r=list(rm), pstar=float(pstar),
                            x=np.array(res.x, dtype=float), fun=float(res.fun),
                            nit=int(res.nit), nfev=int(res.nfev), removed=list(removed),
                            kelly=float(kelly_measure(lm, rm))))
        if stop:
            break
        while not st.notinC.isEmpty():                                             # :1037-1039
            kappa.append(st.notinC.loc[0])
            st.notinC.removeIndex(st.notinC.loc[0])
        random.seed(13)                                                            # :1042-1046
        for i in random.sample(kappa, min(kappasize, len(kappa))):
            st.notinC.insertAngle(i, st.pointtoAngle[i])
            kappa.remove(i)
        if max_iter is not None and len(history) >= max_iter:
            break
    return history


# --------------------------------------------------------------------------
# numpy's float64 add-reduce order (what ``beamGrad[idx].sum()`` does), restated
# --------------------------------------------------------------------------
def numpy_pairwise_sum(a) -> float:
    """Order of numpy's contiguous float64 ``sum`` for n <= 128: sequential below 8
    elements, else 8 interleaved accumulators combined as ((0+1)+(2+3))+((4+5)+(6+7))
    followed by the sequential tail.  The pricing kernel reproduces this order."""
    a = [float(v) for v in a]
    n = len(a)
    if n < 8:
        s = 0.0   # numpy starts from -0.0; identical unless every term is -0.0
        for v in a:
            s = s + v
        return s
    assert n <= 128
    r = a[:8]
    i = 8
    while i < n - (n % 8):
        for k in range(8):
            r[k] = r[k] + a[i + k]
        i += 8
    s = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        s = s + a[i]
        i += 1
    return s


# --------------------------------------------------------------------------
# k-medoid ordering (kmedoids.py:32-68)
# --------------------------------------------------------------------------
def km_distancemin(kappa, points) -> float:
    """Sum over all points of the distance to the nearest medoid (kmedoids.py:32-42).
    `points` is an (n, d) array or a sequence of scalars (the reference's [[0, z]])."""
    P = np.asarray(points, dtype=float)
    if P.ndim == 1:
        P = np.stack([np.zeros_like(P), P], axis=1)
    diff = P[np.asarray(kappa)][:, None, :] - P[None, :, :]
    dist = np.sqrt((diff * diff).sum(axis=2))
    return sum(dist.min(axis=0))


def km_insertnewelement(kappa, points):
    """Append the non-medoid whose insertion minimises km_distancemin; strict ``<`` keeps
    the first best candidate (kmedoids.py:48-60)."""
    n = len(points)
    inset = set(kappa)
    best, flag = None, np.inf
    for c in range(n):
        if c in inset:
            continue
        cost = km_distancemin(kappa + [c], points)
        if cost < flag:
            best, flag = c, cost
    kappa.append(best)
    return kappa


def km_givemewholelist(kappa, points):
    """Greedy k-medoid ordering of all points (kmedoids.py:62-68)."""
    kappa = list(kappa)
    while len(kappa) < len(points):
        kappa = km_insertnewelement(kappa, points)
    return kappa
