#!/usr/bin/env python
"""Stream sizes of a case without a GPU: a Python twin of the plan's layout rules (lanes per row from
the chunk cap, super-slices of W slices, headers) applied to the row lengths of both layouts.
Prints, per layout, nonzero bytes, streamed bytes, header bytes and the padding overhead - the numbers
to look at before changing caps, tile sizes or index widths.  (Estimate: the library is the reference.)"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vmatwpencode_b200.synth import make_config


def stream_bytes(groups, W, cap, header_units, chunk_units):
    units = hdr = nss = 0
    for L in groups:                      # L: lengths sorted descending
        pos = 0
        while pos < len(L):
            nch1 = (int(L[pos]) + 3) // 4
            tpr = 1
            while (nch1 + tpr - 1) // tpr > cap and tpr < 32:
                tpr *= 2
            nch = (nch1 + tpr - 1) // tpr
            units += W * header_units + nch * W * chunk_units
            hdr += W * header_units
            nss += 1
            pos += (32 // tpr) * W
    return units * 16, hdr * 16, nss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--cap1", type=int, default=48)
    ap.add_argument("--cap2", type=int, default=24)
    ap.add_argument("--value-bytes", type=int, default=4)
    ap.add_argument("--index-bytes", type=int, default=2)
    args = ap.parse_args()
    case = make_config(args.workload, args.scale)
    D = case.D.tocsr()
    V, B = D.shape
    T = args.tile or (6144 if V > 500000 else 4096)
    ku = 8 * args.index_bytes + 8 * args.value_bytes
    per_nnz = args.value_bytes + args.index_bytes
    base = 8 if args.index_bytes == 1 else 0
    out = dict(workload=args.workload, V=V, B=B, nnz=int(D.nnz), tile_voxels=T, bytes_per_nonzero=per_nnz)
    L1 = np.sort(np.diff(D.indptr))[::-1]
    b1, h1, n1 = stream_bytes([L1], 16, args.cap1, 8 + 3 * 8 + base, ku)
    out["dose_layout"] = dict(stream_MB=round(b1 / 1e6, 1), header_MB=round(h1 / 1e6, 1), super_slices=n1,
                              overhead_pct=round(100 * (b1 / (D.nnz * per_nnz) - 1), 2))
    Dc = D.tocsc()
    nt = (V + T - 1) // T
    tile_of = Dc.indices // T
    groups = [[] for _ in range(nt)]
    counts = np.zeros((B, nt), dtype=np.int64)
    np.add.at(counts, (np.repeat(np.arange(B), np.diff(Dc.indptr)), tile_of), 1)
    for t in range(nt):
        L = counts[:, t]
        groups[t] = np.sort(L[L > 0])[::-1]
    b2, h2, n2 = stream_bytes(groups, 4, args.cap2, 8 + base, ku)
    seg = np.concatenate(groups)
    out["transposed_layout"] = dict(stream_MB=round(b2 / 1e6, 1), header_MB=round(h2 / 1e6, 1), super_slices=n2,
                                    segments=int(len(seg)), mean_segment=round(float(seg.mean()), 1),
                                    overhead_pct=round(100 * (b2 / (D.nnz * per_nnz) - 1), 2))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
