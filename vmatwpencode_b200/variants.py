"""The two EARLIER generations of the reference's driver functions on the B200 path.

``greedyVMATshort.py`` (first generation; BASELINE.json's configs[0] names it) and ``greedyVMAT.py``
(second) define the same function names as ``greedyVMATsampling.py`` with different signatures and, in
ten places, different semantics.  Two namespaces mirror them:

    from vmatwpencode_b200 import VMATlibrary as vl, variants
    vl.bind(case)                       # uploads D; the evaluation kernels are the same for all generations
    g = variants.short                  # or variants.classic
    g.colGen(C)                         # greedyVMATshort.py:746-814           (classic: colGen(C, WholeCircle, initialApertures), greedyVMAT.py:779-836)
    g.PricingProblem(C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim, N, M)      # :618-667   (classic: :670-700)
    g.PPsubroutine(C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim, predec, succ, N, M, index)   # :456-616 (classic: :469-636)
    g.updateOpenAperture(i) -> (openaperturenp, diagmaker)                                # :832-857   (classic: :856-881)
    g.solveRMC()                                                                          # :669-695   (classic: :702-728)
    g.calcObjGrad(x)                                                                      # :450-454   (classic: :452-456)

``data.notinC`` / ``data.caligraphicC`` are plain lists of control-point indices here, as in those files.
The pricing network runs in the same CUDA kernel as the canonical generation (K5), wired per generation
(``vmat_price_ex``, include/vmat_b200.h): shifted relaxation window, predecessor bias, sink range, untouched
weights 0.0, raw / width-only dose indices, leaf vectors that end with the sink's (0, 0).  The host part
below evaluates the generation's own leaf-range expressions for all rows at once, in float64.
There is no CPU fallback.
"""
from __future__ import annotations

import math
import time
import warnings

import numpy as np
from scipy.optimize import minimize

from . import VMATlibrary as vl
from ._lib import PriceParams, PriceRow, VmatError, VMAT_PRICE_CLASSIC, VMAT_PRICE_SHORT

_ROW = vl._PRICE_ROW_DTYPE


class _Generation:
    def __init__(self, name, code, default_left, ftol, maxiter, vmax, speedlim):
        self.name, self.code, self.default_left = name, code, default_left
        self.ftol, self.maxiter, self.vmax, self.speedlim = ftol, maxiter, vmax, speedlim
        self.gastep = None          # classic: degrees between neighbouring control points (module global `gastep`, greedyVMAT.py:33)

    # -- evaluation ---------------------------------------------------------------------------------
    def calcObjGrad(self, x, user_data=None):
        """(objective, aperture gradient): every open beamlet of a selected control point counts 1
        (greedyVMATshort.py:110-141,450-454; greedyVMAT.py:118-146,452-456)."""
        d = vl.data
        for i in d.caligraphicC:
            if len(d.strengths[i]) != len(d.openApertureMaps[i]):
                d.strengths[i] = [1.0] * len(d.openApertureMaps[i])      # no fractional strengths in these generations
        return vl.calcObjGrad(x)

    # -- pricing --------------------------------------------------------------------------------------
    def _row_plan(self, M):
        if self.name == "short":      # row 0, then m = 2..M-1 with the beamlets of row m-1 and the leaf limits of row m (:499,550-565)
            return [(0, 0)] + [(m, m - 1) for m in range(2, M)]
        return [(m, m) for m in range(0, M - 1)]          # rows 0..M-2 (greedyVMAT.py:510,569)

    def _fill_rows(self, rows, base, am, ap, vmax, speedlim, predec, succ, Nn, Mm, index):
        d = vl.data
        vmaxm = vmaxp = vmax
        if type(predec) is list:
            lcm, rcm, vmaxm = [self.default_left] * Mm, [Nn] * Mm, float("inf")
        else:
            lcm, rcm = d.llist[predec], d.rlist[predec]
        if type(succ) is list:
            lcp, rcp, vmaxp = [self.default_left] * Mm, [Nn] * Mm, float("inf")
        else:
            lcp, rcp = d.llist[succ], d.rlist[succ]
        tm = vmaxm * am / speedlim
        tp = vmaxp * ap / speedlim
        plan = self._row_plan(Mm)
        out = np.frombuffer(rows, dtype=_ROW)[base:base + len(plan)]
        grid = vl._beamlet_grid(index)
        leftmost = 0
        for k, (row, brow) in enumerate(plan):
            vb = np.flatnonzero(grid[brow])
            if vb.size == 0:
                raise ValueError("min() arg is an empty sequence")
            vmin, vmx, cnt = int(vb[0]), int(vb[-1]), len(vb)
            o = out[k]
            o["n_r_mid"] = 0
            o["r_mid"] = 0.0
            if self.name == "short":
                lo = math.ceil(max(vmin - 1, lcm[row] - tm, lcp[row] - tp))
                hi = math.floor(min(vmx, lcm[row] + tm, lcp[row] + tp))                  # exclusive
                o["l_first"], o["n_l"] = lo, max(0, hi - lo)
                o["r_lo"] = max(rcm[row] - tm, rcp[row] - tp)
                o["r_hi"] = min(vmx + 1, rcm[row] + tm, rcp[row] + tp) - 1               # range(..., floor(x)) is exclusive
                if k == 0:          # beamGrad[l+1:r]: raw positions
                    o["vb_mask"], o["vb_min"], o["grad_offset"] = np.uint64((1 << min(Nn + 1, 63)) - 1), 0, 0
                else:               # beamGrad[leftmost+1 : (r-l)+leftmost]
                    o["vb_mask"], o["vb_min"], o["grad_offset"] = np.uint64(0), 0, leftmost
                leftmost = cnt if k == 0 else leftmost + cnt
            else:
                lo = math.ceil(max(-1, lcm[row] - tm, lcp[row] - tp))
                hi = 1 + math.floor(min(Nn - 1, lcm[row] + tm, lcp[row] + tp))
                r_lo = max(rcm[row] - tm, rcp[row] - tp)
                r_hi = min(Nn, rcm[row] + tm, rcp[row] + tp)
                if hi - lo <= 0:
                    if k == 0:
                        raise NameError("name 'm' is not defined")     # the reference's message reads m before the row loop (:533)
                    hi = lo + 1
                if k == 0:          # an empty r range in row 0 runs into the same message
                    for l in range(lo, hi):
                        if math.floor(r_hi) < math.ceil(max(l + 1, r_lo)):
                            raise NameError("name 'm' is not defined")
                o["l_first"], o["n_l"] = lo, hi - lo
                o["r_lo"], o["r_hi"] = r_lo, r_hi
                o["r_mid"] = lo                                        # empty r range: r0 .. the row's first l (:539-540,583-584)
                o["vb_mask"] = np.uint64(sum(1 << int(y) for y in vb))
                o["vb_min"] = vmin
                o["grad_offset"] = 0 if k == 0 else leftmost
                leftmost = cnt - 1 if k == 0 else leftmost + cnt
        return len(plan)

    @staticmethod
    def _node_counts(arr, classic):
        """Nodes of every described row (the kernel's enumeration, on the host)."""
        out = np.zeros(len(arr), dtype=np.int64)
        for k, o in enumerate(arr):
            tot = 0
            r1 = math.floor(o["r_hi"])
            for q in range(int(o["n_l"])):
                r0 = math.ceil(max(o["l_first"] + q + 1.0, o["r_lo"]))
                tot += (r1 - r0 + 1) if r1 >= r0 else (max(0, int(o["r_mid"] + 1.0 - r0)) if classic else 0)
            out[k] = tot
        return out

    def _price(self, cands, C, C2, C3, b, vmax, speedlim, Nn, Mm):
        d = vl.data
        if d._fg is None:
            raise VmatError("pricing needs the beamlet gradient of an evaluation: call data.calcDose() first")
        if not d._gb_valid:
            d.calcDose()
        nrows = len(self._row_plan(Mm))
        rows = (PriceRow * (len(cands) * nrows))()
        for k, (idx, predec, succ, am, ap) in enumerate(cands):
            self._fill_rows(rows, k * nrows, am, ap, vmax, speedlim, predec, succ, Nn, Mm, idx)
        arr = np.frombuffer(rows, dtype=_ROW)
        classic = self.name == "classic"
        counts = self._node_counts(arr, classic).reshape(len(cands), nrows)
        # Rows without nodes (leaf speeds so tight that no position is reachable).  The reference then either runs
        # np.argmin over an empty window in the next row that has nodes (ValueError), or - when no later row has
        # nodes - reaches its sink loop, which covers nothing (short: p stays unassigned) or only the last node
        # of the last row that had any (classic).
        used = np.full(len(cands), nrows)
        for k in range(len(cands)):
            empty = np.flatnonzero(counts[k] == 0)
            if empty.size:
                first = int(empty[0])
                if (counts[k, first:] > 0).any() or first == 0:
                    raise ValueError("attempt to get argmin of an empty sequence")
                if not classic:
                    raise UnboundLocalError("local variable 'p' referenced before assignment")
                used[k] = first
        p = np.empty(len(cands))
        l, r = [None] * len(cands), [None] * len(cands)
        for m_used in sorted(set(used.tolist())):             # normally one group: every row has nodes
            grp = [k for k in range(len(cands)) if used[k] == m_used]
            sub = (PriceRow * (len(grp) * m_used))()
            sub_arr = np.frombuffer(sub, dtype=_ROW)
            for q, k in enumerate(grp):
                sub_arr[q * m_used:(q + 1) * m_used] = arr[k * nrows:k * nrows + m_used]
            prm = PriceParams(C=float(C), C2=float(C2), C3=float(C3), b=float(b), M=m_used, N=Nn, variant=self.code)
            prm.reserved[0] = 1 if m_used < nrows else 0
            pp, ll, rr = d.plan.price_ex([cands[k][0] for k in grp], prm, sub)
            for q, k in enumerate(grp):
                p[k], l[k], r[k] = pp[q], ll[q], rr[q]
        if np.isnan(p).any():
            raise UnboundLocalError("local variable 'p' referenced before assignment")
        return p, l, r

    def PPsubroutine(self, C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim, predec, succ, N, M, index):
        p, l, r = self._price([(index, predec, succ, angdistancem, angdistancep)], C, C2, C3, b, vmax, speedlim, N, M)
        return np.float64(p[0]), [np.int64(v) for v in l[0]], [np.int64(v) for v in r[0]]

    def _neighbours(self, idx):
        d = vl.data
        if self.name == "short":      # range() over a list raises inside a bare try: never any neighbour (:635-651)
            return [], []
        succs = [i for i in d.caligraphicC if i > idx]
        predecs = [i for i in d.caligraphicC if i < idx]
        return (max(predecs) if predecs else []), (min(succs) if succs else [])

    def PricingProblem(self, C, C2, C3, b, *args):
        """short: PricingProblem(C, C2, C3, b, angdistancem, angdistancep, vmax, speedlim, N, M);
        classic: PricingProblem(C, C2, C3, b, vmax, speedlim, N, M).  -> (pstar, l, r, bestAperture)"""
        d = vl.data
        if self.name == "short":
            am, ap, vmax, speedlim, Nn, Mm = args
        else:
            vmax, speedlim, Nn, Mm = args
        cands = []
        for idx in d.notinC:
            predec, succ = self._neighbours(idx)
            if self.name != "short":
                if self.gastep is None:
                    raise VmatError("variants.classic.gastep (degrees between control points) is not set")
                ap = (succ - idx) * self.gastep if type(succ) is not list else np.inf
                am = (idx - predec) * self.gastep if type(predec) is not list else np.inf
            cands.append((idx, predec, succ, am, ap))
        if not cands:
            if self.name == "short":
                return float("inf"), [], [], None
            raise ValueError("attempt to get argmin of an empty sequence")
        p, l, r = self._price(cands, C, C2, C3, b, vmax, speedlim, Nn, Mm)
        k = int(np.argmin(p))                 # first minimum = the strict `<` scan of greedyVMATshort.py:654
        return (np.float64(p[k]), [np.int64(v) for v in l[k]], [np.int64(v) for v in r[k]], cands[k][0])

    # -- aperture bookkeeping ---------------------------------------------------------------------------
    def updateOpenAperture(self, i):
        """(open beamlet indices, diagmaker): beamlets l+1 .. r-1 of every row (greedyVMATshort.py:832-857 =
        greedyVMAT.py:856-881)."""
        d = vl.data
        seen = 0
        open_idx = []
        grid = vl._beamlet_grid(i)
        for m in range(len(d.llist[i])):
            vb = np.flatnonzero(grid[m])
            lo = d.llist[i][m] - int(vb.min()) + 1 + seen
            hi = d.rlist[i][m] - int(vb.min()) - 1 + seen
            seen += len(vb)
            if lo < hi + 1:
                open_idx.extend(range(int(lo), int(hi) + 1))
        oam = np.array(open_idx, dtype=int)
        diag = np.zeros(int(d.beam_ptr[i + 1] - d.beam_ptr[i]), dtype=float)
        diag[[k for k in oam]] = 1.0
        return oam, diag

    def solveRMC(self):
        """L-BFGS-B over all control points, free where selected, no upper bound (greedyVMATshort.py:669-695,
        greedyVMAT.py:702-728)."""
        d = vl.data
        start = time.time()
        self.calcObjGrad(d.currentIntensities)
        bounds = [(0, 0) if k in d.notinC else (0, None) for k in range(d.numbeams)]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = minimize(self.calcObjGrad, d.currentIntensities, method="L-BFGS-B", jac=True, bounds=bounds,
                           options={"ftol": self.ftol, "maxiter": self.maxiter})
        d.rmp_seconds = time.time() - start
        return res

    def colGen(self, C, WholeCircle=False, initialApertures=None, max_iter=None):
        """The greedy loop (greedyVMATshort.py:746-814: colGen(C); greedyVMAT.py:779-836: colGen(C, WholeCircle,
        initialApertures), whose WholeCircle branch cannot run in the reference - ``kappa.pop(range(...))``).
        Steps are appended to data.history; returns pstar."""
        d = vl.data
        if WholeCircle:
            raise TypeError("'range' object cannot be interpreted as an integer")      # what kappa.pop(range(...)) raises there (:800)
        C2 = C3 = 1.0
        Nn, Mm = len(d.yinter), len(d.xinter)
        d.llist = [[-1] * Mm for _ in range(d.numbeams)]
        d.rlist = [[Nn + 1] * Mm for _ in range(d.numbeams)]
        d.strengths = [[] for _ in range(d.numbeams)]
        d.caligraphicC = []
        d.notinC = list(range(d.numbeams))
        d.calcDose()
        pstar = -float("inf")
        d.history = []
        while pstar < 0 and (self.name == "short" or len(d.notinC) > 0):
            d.calcDose()
            d.calcGradientandObjValue()
            if self.name == "short":
                pstar, lm, rm, best = self.PricingProblem(C, C2, C3, 0.5, 60, 60, self.vmax, self.speedlim, Nn, Mm)
            else:
                pstar, lm, rm, best = self.PricingProblem(C, C2, C3, 0.5, self.vmax, self.speedlim, Nn, Mm)
            if pstar >= 0:
                break
            d.caligraphicC.append(best)
            d.caligraphicC.sort()
            d.notinC.remove(best)
            d.notinC.sort()
            d.llist[best], d.rlist[best] = lm, rm
            d.openApertureMaps[best], d.diagmakers[best] = self.updateOpenAperture(best)
            d.strengths[best] = [1.0] * len(d.openApertureMaps[best])
            res = self.solveRMC()
            d.rmpres = res
            d.history.append(dict(idx=int(best), l=[int(v) for v in lm], r=[int(v) for v in rm], pstar=float(pstar),
                                  x=np.array(res.x, dtype=float), fun=float(res.fun), nit=int(res.nit)))
            if max_iter is not None and len(d.history) >= max_iter:
                break
        return pstar


short = _Generation("short", VMAT_PRICE_SHORT, 0, 1e-6, 1000, 2.0, 3.0)          # greedyVMATshort.py:481,686-692,748-753
classic = _Generation("classic", VMAT_PRICE_CLASSIC, -1, 1e-2, 15, 2.0 * 1000, 3.0)   # greedyVMAT.py:492,719-725,783-786
