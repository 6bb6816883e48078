#!/usr/bin/env python
"""Stream sizes of a case without a GPU: a Python twin of the plan's layout rules (lanes per row from
the chunk cap, super-slices of W slices, headers) applied to the row lengths of both layouts.
Prints, per layout, nonzero bytes, streamed bytes, header bytes and the padding overhead (with --index-bytes 1 also the
bridge entries of the byte-delta layout) - the numbers to look at before changing caps, tile sizes or index widths.  (Estimate: the library is the reference.)"""
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


def bridges(index, segment, chain255=False):
    """Bridge entries in front of every nonzero (byte-delta layout): ceil((gap >> 8) / 255) zero-valued entries whose
    delta counts units of 256 (bridge_entries() in vmat_eval_kernels.cuh); chain255: the first version of the layout,
    one entry per 255 of the gap.  `segment` tells rows / (beamlet, tile) segments apart; explicit zeros are ignored."""
    gap = np.diff(index.astype(np.int64), prepend=0)
    same = np.concatenate([[False], segment[1:] == segment[:-1]])
    gap = np.where(same, gap, 0)
    if chain255:
        return np.where(gap > 255, (gap - 1) // 255, 0)
    return ((gap >> 8) + 254) // 255


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
    delta = args.index_bytes == 1
    D.sort_indices()
    rows1 = np.repeat(np.arange(V), np.diff(D.indptr))
    L1 = np.diff(D.indptr).astype(np.int64)
    if delta:                             # byte deltas: slots = nonzeros + bridge entries (k_count_pads)
        br1 = bridges(D.indices, rows1)
        L1 = L1 + np.bincount(rows1, weights=br1, minlength=V).astype(np.int64)
        out["dose_bridge_entries_pct"] = round(100 * float(br1.sum()) / D.nnz, 2)
        out["dose_bridge_entries_pct_255_chain"] = round(100 * float(bridges(D.indices, rows1, chain255=True).sum()) / D.nnz, 2)
    L1 = np.sort(L1)[::-1]
    b1, h1, n1 = stream_bytes([L1], 16, args.cap1, 8 + 3 * 8 + base, ku)
    out["dose_layout"] = dict(stream_MB=round(b1 / 1e6, 1), header_MB=round(h1 / 1e6, 1), super_slices=n1,
                              overhead_pct=round(100 * (b1 / (D.nnz * per_nnz) - 1), 2))
    Dc = D.tocsc()
    Dc.sort_indices()
    nt = (V + T - 1) // T
    tile_of = Dc.indices // T
    groups = [[] for _ in range(nt)]
    counts = np.zeros((B, nt), dtype=np.int64)
    cols2 = np.repeat(np.arange(B), np.diff(Dc.indptr))
    np.add.at(counts, (cols2, tile_of), 1)
    if delta:
        br2 = bridges(Dc.indices, cols2.astype(np.int64) * nt + tile_of)      # gaps inside a (beamlet, tile) segment
        np.add.at(counts, (cols2, tile_of), br2.astype(np.int64))
        out["transposed_bridge_entries_pct"] = round(100 * float(br2.sum()) / D.nnz, 2)
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
