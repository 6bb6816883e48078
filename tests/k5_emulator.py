"""Host twin of the pricing kernel (csrc/vmat_price_kernels.cuh): consumes the same per-row
descriptors the product builds on the host and runs the same node enumeration, cost formulas
(float64, one rounding per operation), first/last-minimum rules and backtrack.  Lets the
descriptor logic and the kernel's algorithm be checked against the oracle without a GPU."""
import math

import numpy as np

from oracle.vmat_oracle import numpy_pairwise_sum


def nodes_from_row(row):
    out = []
    for k in range(row.n_l):
        l = row.l_first + k
        r0 = math.ceil(max(l + 1.0, row.r_lo))
        r1 = math.floor(row.r_hi)
        if r1 >= r0:
            out.append((l, [float(r) for r in range(r0, r1 + 1)]))
        else:
            out.append((l, [row.r_mid + k2 for k2 in range(row.n_r_mid)]))
    return out


def emulate_price(rows, M, C, C2, C3, b, bg):
    """rows: M descriptors of one candidate; bg: that control point's beamlet gradient."""
    f = np.float64
    C, C2, C3b = f(C), f(C2), f(C3) * f(b)
    nbl = len(bg)
    prev = None
    hist = []
    for m in range(M):
        R = rows[m]
        cur = []
        for l, rs in nodes_from_row(R):
            for r in rs:
                l_, r_ = f(l), f(r)
                lp1 = l_ + f(1.0)
                y_lo, y_hi = int(math.ceil(lp1)), int(math.floor(r_))
                sel = [y for y in range(max(y_lo, 0), min(y_hi, 64)) if (R.vb_mask >> y) & 1]
                width = r_ - l_
                i_l = min(max(int(math.floor(lp1)), 0), nbl - 1)
                i_r = min(max(int(math.ceil(r_)), 0), nbl - 1)
                side = -((f(math.ceil(lp1)) - lp1) * bg[i_l] + (r_ - f(math.floor(r_))) * bg[i_r])
                dose = f(0.0)
                if sel:
                    dose = -f(numpy_pairwise_sum([bg[R.grad_offset - R.vb_min + y] for y in sel]))
                tie = f(10E-10) * width
                node = dict(lf=l_, rf=r_, li=int(l_), ri=int(r_), dose=dose, tie=tie, side=side,
                            c3s=(C3b * width) if sel else f(0.0), has=bool(sel))
                if m == 0:
                    w = f(0.0)
                    if lp1 < r_:
                        pen = C * (C2 * width - C3b * width)
                        w = ((pen - dose) + tie) + side if sel else (pen + tie) + side
                    node["w"], node["dad"] = w, -1
                else:
                    best, bj = None, -1
                    for j, pv in enumerate(prev):
                        pl, pr = f(pv["li"]), f(pv["ri"])
                        lam = abs(pl - l_) + abs(pr - r_)
                        lam = lam - f(2.0) * max(f(0.0), pl - r_)
                        lam = lam - f(2.0) * max(f(0.0), l_ - abs(pr))
                        wgt = C * (C2 * lam - node["c3s"])
                        wgt = ((wgt - dose) + tie) + side
                        tot = pv["w"] + wgt
                        if best is None or tot < best:
                            best, bj = tot, j
                    node["w"], node["dad"] = best, bj
                cur.append(node)
        hist.append(cur)
        prev = cur
    best, bk = None, -1
    for k, nd in enumerate(prev):
        tot = nd["w"] + C * (C2 * f(nd["ri"] - nd["li"]))
        if best is None or tot <= best:
            best, bk = tot, k
    l_out, r_out = [0] * M, [0] * M
    k = bk
    for m in range(M - 1, -1, -1):
        l_out[m], r_out[m] = hist[m][k]["li"], hist[m][k]["ri"]
        k = hist[m][k]["dad"]
    return float(best), l_out, r_out
