"""TEST INFRASTRUCTURE - CPU restatement of the pricing network and driver functions of the two EARLIER
generations of the reference: greedyVMATshort.py ("short") and greedyVMAT.py ("classic").

SURVEY.md section 2.2 tabulates how they differ from the canonical greedyVMATsampling.py restated in
oracle/vmat_oracle.py; this module parametrises on exactly those rows.  Like the reference, the network
is held in global node arrays (node 0 = source, rows consecutive from node 1), because two of the
differences only exist in that numbering: the relaxation window of a row is the previous row SHIFTED BY
ONE node (it starts at the node before the row's first one and drops its last one), and greedyVMAT.py
records as predecessor the node AFTER the one whose weight it used.

Parity pinning: the reference's own functions, AST-extracted (oracle/ref_extract.py, modules
"greedyVMATshort" / "greedyVMAT"), live in tests/test_oracle_vs_reference.py and frozen by
oracle/make_golden.py into tests/golden/variant_*.npz.  Not product code.
"""
from __future__ import annotations

import math
import warnings
from dataclasses import dataclass

import numpy as np
from scipy.optimize import minimize

from . import vmat_oracle as O


@dataclass(frozen=True)
class Variant:
    """One column of SURVEY.md 2.2."""
    name: str
    default_left: int        # predecessor / successor left leaves when there is none (short 0, classic -1)
    window_shift: int        # 1: relax over [node before the previous row's first, ..., its last but one]
    dad_bias: int            # added to the window position when the predecessor is recorded
    sink_extend: int         # 1: the sink also scans the node after its window (= the previous row's last)
    tie_break: bool          # + 10E-10 * (r - l)
    rows: str                # "short": row 0, then m = 2..M-1 with the beamlets of row m-1; "classic": rows 0..M-2
    ftol: float
    maxiter: int
    vmax: float
    speedlim: float


SHORT = Variant("short", 0, 1, 0, 0, False, "short", 1e-6, 1000, 2.0, 3.0)        # greedyVMATshort.py:456-616,686-692,748-753
CLASSIC = Variant("classic", -1, 1, 1, 1, True, "classic", 1e-2, 15, 2000.0, 3.0)  # greedyVMAT.py:469-636,719-725,783-786
VARIANTS = {"short": SHORT, "classic": CLASSIC}


class VariantState(O.OracleState):
    """OracleState with the plain-list candidate sets these generations use (data.notinC / caligraphicC are
    lists of control-point indices, greedyVMATshort.py:766-767, greedyVMAT.py:796-801)."""

    def __init__(self, case):
        super().__init__(case)
        self.notinC = []
        self.caligraphicC = []


# --------------------------------------------------------------------------
# evaluation: no strengths in these generations (greedyVMATshort.py:110-141, greedyVMAT.py:118-146)
# --------------------------------------------------------------------------
def calc_obj_grad(st: VariantState, x):
    """calcObjGrad of both generations in their own operation order: every open beamlet counts 1
    (CSC column slice x repeat(y_i)), dZdK column = row sums of D_i^T diag(diagmaker), dense V x nb product."""
    from scipy import sparse
    st.currentIntensities = x
    dose = np.zeros(st.numvoxels, dtype=float)
    dZdK = np.zeros((st.numvoxels, st.numbeams))
    for i in st.caligraphicC:
        oam = st.openApertureMaps[i]
        dose += st.DlistT[i][:, oam] @ np.repeat(st.currentIntensities[i], len(oam))
        col = (st.DlistT[i] @ sparse.diags(np.asarray(st.diagmakers[i], dtype=float), 0)).sum(axis=1)
        dZdK[:, i] = np.asarray(col).ravel()
    st.currentDose = dose
    st.objectiveValue, st.voxelgradient = O.penalty(dose, st.thr, st.over, st.under)
    st.aperturegradient = (st.voxelgradient[None, :] @ dZdK).T
    return st.objectiveValue, st.aperturegradient


# --------------------------------------------------------------------------
# pricing network
# --------------------------------------------------------------------------
def _row_plan(var: Variant, M):
    """(leaf-limit row, beamlet row) of every processed network row."""
    if var.rows == "short":       # greedyVMATshort.py:499,550-565: row 0, then m = 2..M-1 reading xinter[m-1] but lcm[m]
        return [(0, 0)] + [(m, m - 1) for m in range(2, M)]
    return [(m, m) for m in range(0, M - 1)]          # greedyVMAT.py:510,569: rows 0..M-2


def _ranges(var: Variant, row, vb, N, lcm, lcp, rcm, rcp, tm, tp, network_row):
    """[(l, [r...])] of one row (greedyVMATshort.py:527-528,565-566; greedyVMAT.py:529-545,576-586)."""
    out = []
    if var.rows == "short":
        ls = range(math.ceil(max(min(vb) - 1, lcm[row] - tm, lcp[row] - tp)),
                   math.floor(min(max(vb), lcm[row] + tm, lcp[row] + tp)))
        for l in ls:
            rs = range(math.ceil(max(l + 1, rcm[row] - tm, rcp[row] - tp)),
                       math.floor(min(max(vb) + 1, rcm[row] + tm, rcp[row] + tp)))
            out.append((l, list(rs)))
        return out
    ls = range(math.ceil(max(-1, lcm[row] - tm, lcp[row] - tp)), 1 + math.floor(min(N - 1, lcm[row] + tm, lcp[row] + tp)))
    if len(ls) == 0:
        if network_row == 0:
            raise NameError("name 'm' is not defined")       # the reference's message print reads m before the row loop (:533)
        ls = range(ls.start, ls.start + 1)
    for l in ls:
        rs = range(math.ceil(max(l + 1, rcm[row] - tm, rcp[row] - tp)), 1 + math.floor(min(N, rcm[row] + tm, rcp[row] + tp)))
        if len(rs) == 0:
            if network_row == 0:
                raise NameError("name 'm' is not defined")
            rs = range(rs.start, ls.start + 1)
        out.append((l, list(rs)))
    return out


def pp_subroutine(st: VariantState, var: Variant, C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim,
                  predec, succ, N, M, index, beamGrad=None):
    """PPsubroutine of greedyVMATshort.py:456-616 / greedyVMAT.py:469-636 -> (p, l, r); l and r end with
    the sink's (0, 0) entry, as there.  Raises UnboundLocalError where the reference does (no path into
    the sink that is <= the sink's initial 0.0)."""
    vmaxm = vmaxp = vmax
    if isinstance(predec, list):
        lcm, rcm, vmaxm = [var.default_left] * M, [N] * M, float("inf")
    else:
        lcm, rcm = st.llist[predec], st.rlist[predec]
    if isinstance(succ, list):
        lcp, rcp, vmaxp = [var.default_left] * M, [N] * M, float("inf")
    else:
        lcp, rcp = st.llist[succ], st.rlist[succ]
    if beamGrad is None:
        beamGrad = st.Dlist[index] @ st.voxelgradient
    tm = vmaxm * angdistancem / speedlim
    tp = vmaxp * angdistancep / speedlim
    cap = 50 * 50 * (M + 2)                                   # the reference's over-estimate of the node count
    lnet = np.zeros(cap, dtype=np.int64)
    rnet = np.zeros(cap, dtype=np.int64)
    wnet = np.zeros(cap, dtype=float)                          # zeros: node 0 and the sink start at 0.0
    dad = np.zeros(cap, dtype=np.int64)
    node = 0
    pos = 0            # posBeginningOfRow
    nprev = 0
    leftmost = 0
    for k, (row, brow) in enumerate(_row_plan(var, M)):
        vb = O.valid_beamlets(st, brow, index)
        old, nprev = nprev, 0
        for l, rs in _ranges(var, row, vb, N, lcm, lcp, rcm, rcp, tm, tp, k):
            for r in rs:
                node += 1
                nprev += 1
                lnet[node], rnet[node] = l, r
                if k == 0:
                    if var.rows == "short":                    # :530-536 raw slice, no tie-break
                        if l + 1 <= r - 1:
                            dose = -beamGrad[l + 1:r].sum()
                            w = C * (C2 * (r - l) - C3 * b * (r - l)) - dose
                        else:
                            w = 0.0
                    else:                                      # :546-556
                        w = 0.0
                        if l + 1 < r:
                            idx = np.intersect1d(range(l + 1, r), vb) - min(vb)
                            if len(idx) > 0:
                                dose = -beamGrad[idx].sum()
                                w = C * (C2 * (r - l) - C3 * b * (r - l)) - dose + 10E-10 * (r - l)
                    wnet[node] = w
                    continue
                if var.rows == "short":                        # :574-582 width-only slice
                    lo, hi = leftmost, (r - l) + leftmost
                    if lo + 1 <= hi - 1:
                        dose, c3s = -beamGrad[lo + 1:hi].sum(), C3 * b * (r - l)
                    else:
                        dose, c3s = 0.0, 0
                else:                                          # :594-601
                    idx = np.intersect1d(range(l + 1, r), vb) + leftmost - min(vb)
                    if len(idx) > 0:
                        dose, c3s = -beamGrad[idx].sum(), C3 * b * (r - l)
                    else:
                        dose, c3s = 0.0, 0.0
                win = slice(pos - old, pos)                    # shifted by one node against the previous row
                lw, rw = lnet[win], rnet[win]
                lam = (np.absolute(lw - l) + np.absolute(rw - r) - 2 * np.maximum(0, lw - r)
                       - 2 * np.maximum(0, l - np.absolute(rw)))
                wgt = C * (C2 * lam - c3s) - dose
                if var.tie_break:
                    wgt = wgt + 10E-10 * (r - l)
                tot = wnet[win] + wgt
                j = int(np.argmin(tot))
                wnet[node] = tot[j]
                dad[node] = j + pos - old + var.dad_bias
        pos += nprev
        if k == 0:
            leftmost = len(vb) if var.rows == "short" else len(vb) - 1
        else:
            leftmost += len(vb)
    node += 1                                                  # the sink
    p = None
    for k in range(pos - nprev, pos + var.sink_extend):
        w = C * (C2 * (rnet[k] - lnet[k]))
        if wnet[k] + w <= wnet[node]:
            wnet[node] = wnet[k] + w
            dad[node] = k
            p = wnet[node]
    l_out, r_out = [], []
    t = node
    while True:
        l_out.append(int(lnet[t]))
        r_out.append(int(rnet[t]))
        t = int(dad[t])
        if t == 0:
            break
    l_out.reverse()
    r_out.reverse()
    if p is None:
        raise UnboundLocalError("local variable 'p' referenced before assignment")
    return p, l_out, r_out


def neighbours(st: VariantState, var: Variant, idx, gastep):
    """greedyVMAT.py:638-664 (selected neighbours by index, gaps = index difference * gastep);
    greedyVMATshort.py:635-651 never finds any (range() over a list raises inside a bare try)."""
    if var.rows == "short":
        return [], []
    succs = [i for i in st.caligraphicC if i > idx]
    predecs = [i for i in st.caligraphicC if i < idx]
    succ = min(succs) if succs else []
    predec = max(predecs) if predecs else []
    return predec, succ


def pricing_problem(st: VariantState, var: Variant, C, C2, C3, b, vmax, speedlim, N, M, gastep=None,
                    angdistancem=60, angdistancep=60):
    """PricingProblem: greedyVMATshort.py:618-667 (serial, strict `<`: first best) /
    greedyVMAT.py:670-700 (np.argmin over the pool's results).  -> (pstar, l, r, best, all)"""
    res = []
    for idx in st.notinC:
        predec, succ = neighbours(st, var, idx, gastep)
        if var.rows == "short":
            am, ap = angdistancem, angdistancep
        else:
            ap = (succ - idx) * gastep if not isinstance(succ, list) else np.inf
            am = (idx - predec) * gastep if not isinstance(predec, list) else np.inf
        p, l, r = pp_subroutine(st, var, C, C2, C3, b, am, ap, vmax, speedlim, predec, succ, N, M, idx)
        res.append((p, l, r, idx))
    if not res:
        return float("inf"), [], [], None, res
    k = int(np.argmin(np.array([q[0] for q in res])))          # first minimum: same as the strict `<` scan
    return res[k][0], res[k][1], res[k][2], res[k][3], res


def update_open_aperture(st: VariantState, i):
    """greedyVMATshort.py:832-857 = greedyVMAT.py:856-881: open beamlets l+1 .. r-1 of every row."""
    seen = 0
    open_idx = []
    for m in range(len(st.llist[i])):
        vb = O.valid_beamlets(st, m, i)
        lo = st.llist[i][m] - min(vb) + 1 + seen
        hi = st.rlist[i][m] - min(vb) - 1 + seen
        seen += len(vb)
        if lo < hi + 1:
            open_idx.extend(range(lo, hi + 1))
    oam = np.array(open_idx, dtype=int)
    diag = np.zeros(st.Dlist[i].shape[0], dtype=float)
    diag[[k for k in oam]] = 1.0
    return oam, diag


def solve_rmc(st: VariantState, var: Variant, evaluator=None):
    """solveRMC: greedyVMATshort.py:669-695 / greedyVMAT.py:702-728 (free where selected, no upper bound)."""
    fn = evaluator or (lambda x: calc_obj_grad(st, x))
    fn(st.currentIntensities)
    bounds = [(0, 0) if k in st.notinC else (0, None) for k in range(st.numbeams)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return minimize(fn, st.currentIntensities, method="L-BFGS-B", jac=True, bounds=bounds,
                        options={"ftol": var.ftol, "maxiter": var.maxiter})


def col_gen(st: VariantState, var: Variant, C, gastep=None, max_iter=None, C2=1.0, C3=1.0, b=0.5):
    """colGen: greedyVMATshort.py:746-814 (colGen(C)) / greedyVMAT.py:779-836 (WholeCircle=False branch)."""
    N, M = st.N, st.M
    st.llist = [[-1] * M for _ in range(st.numbeams)]
    st.rlist = [[N + 1] * M for _ in range(st.numbeams)]
    st.caligraphicC = []
    st.notinC = list(range(st.numbeams))
    calc_obj_grad(st, st.currentIntensities)
    history = []
    pstar = -float("inf")
    while pstar < 0 and (var.rows == "short" or len(st.notinC) > 0):
        calc_obj_grad(st, st.currentIntensities)
        pstar, lm, rm, best, _ = pricing_problem(st, var, C, C2, C3, b, var.vmax, var.speedlim, N, M, gastep)
        if pstar >= 0:
            break
        st.caligraphicC.append(best)
        st.caligraphicC.sort()
        st.notinC.remove(best)
        st.notinC.sort()
        st.llist[best], st.rlist[best] = lm, rm
        st.openApertureMaps[best], st.diagmakers[best] = update_open_aperture(st, best)
        res = solve_rmc(st, var)
        history.append(dict(idx=int(best), l=[int(v) for v in lm], r=[int(v) for v in rm], pstar=float(pstar),
                            x=np.array(res.x, dtype=float), fun=float(res.fun), nit=int(res.nit)))
        if max_iter is not None and len(history) >= max_iter:
            break
    return history
