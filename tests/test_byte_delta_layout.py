"""The byte-delta index layout (index_bytes=1; k_fill_stream / chunk_load in
vmatwpencode_b200/csrc/vmat_eval_kernels.cuh) restated in Python: how a row is expanded with
bridge entries, cut into per-lane blocks and decoded again.  CPU only: pins the arithmetic
(pads per gap, delta range, coverage) the CUDA code implements; the GPU parity tests check the
kernels themselves."""
import numpy as np


def count_pads(cols):
    """k_count_pads: one bridge entry per 255 of every index gap above 255."""
    return int(sum((g - 1) // 255 for g in np.diff(cols) if g > 255))


def lanes_and_chunks(slots, cap):
    """StreamDesc::add_group: lanes per row (power of two) and chunks per lane."""
    nch1 = (slots + 3) // 4
    tpr = 1
    while (nch1 + tpr - 1) // tpr > cap and tpr < 32:
        tpr *= 2
    return tpr, (nch1 + tpr - 1) // tpr


def fill(cols, vals, tpr, nch):
    """k_fill_stream, IW == 1: per lane (start index, [(delta, value)] * 4*nch)."""
    out = []
    for part in range(tpr):
        first, last = part * 4 * nch, (part + 1) * 4 * nch
        e = pads_left = idx = base = prev = 0
        entries = []
        for s in range(last):
            v = 0.0
            if e < len(cols):
                if pads_left > 0:
                    idx += 255
                    pads_left -= 1
                else:
                    idx, v = int(cols[e]), float(vals[e])
                    e += 1
                    if e < len(cols):
                        gap = int(cols[e]) - idx
                        pads_left = (gap - 1) // 255 if gap > 255 else 0
            if s < first:
                continue
            if s == first:
                base = prev = idx
            entries.append((idx - prev, v))
            prev = idx
        out.append((base, entries))
    return out


def decode(lanes):
    """chunk_load, IW == 1: running index per lane; returns {index: summed value}."""
    acc = {}
    for base, entries in lanes:
        cur = base
        for d, v in entries:
            assert 0 <= d <= 255
            cur += d
            if v != 0.0:
                acc[cur] = acc.get(cur, 0.0) + v
    return acc


def test_rows_survive_expansion_blocking_and_decoding():
    rng = np.random.default_rng(0)
    for trial in range(600):
        B = int(rng.choice([300, 5000, 70000]))
        L = int(rng.integers(0, 60))
        cols = np.sort(rng.choice(B, size=min(L, B), replace=False)) if L else np.array([], dtype=int)
        vals = rng.random(len(cols)) + 0.1
        slots = len(cols) + count_pads(cols)
        tpr, nch = lanes_and_chunks(slots, int(rng.choice([1, 2, 5, 24, 48])))
        assert tpr * nch * 4 >= slots                          # every expanded entry has a place
        got = decode(fill(cols, vals, tpr, nch))
        assert got == {int(c): float(v) for c, v in zip(cols, vals)}


def test_gap_edges():
    for gap, pads in ((1, 0), (255, 0), (256, 1), (510, 1), (511, 2), (765, 2), (766, 3)):
        cols = np.array([7, 7 + gap])
        assert count_pads(cols) == pads
        tpr, nch = lanes_and_chunks(2 + pads, 48)
        lanes = fill(cols, [1.0, 2.0], tpr, nch)
        assert decode(lanes) == {7: 1.0, 7 + gap: 2.0}
        assert max(d for _, ent in lanes for d, _ in ent) <= 255
