"""Host twin of the pricing kernel's node enumeration (csrc/vmat_price_kernels.cuh): turns
the per-row descriptors the product builds on the host into the (l, r) node lists the kernel
will visit, so the descriptor logic can be checked against the oracle without a GPU."""
import math


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
