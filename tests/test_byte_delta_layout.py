"""The byte-delta index layout (index_bytes=1; k_count_pads / k_fill_stream / chunk_load in
vmatwpencode_b200/csrc/vmat_eval_kernels.cuh) restated in Python: how a row is expanded with
bridge entries, cut into per-lane blocks and decoded again.  CPU only: pins the arithmetic
(bridges per gap, delta range, zero-valued entries, coverage) the CUDA code implements; the GPU
parity tests check the kernels themselves.

Rule: an entry whose STORED value is zero contributes nothing and its delta byte counts units of
256; every other entry's delta byte counts single indices."""
import numpy as np


def bridge_entries(gap):
    """bridge_entries(): zero-valued entries in front of a nonzero `gap` past its predecessor."""
    return ((gap >> 8) + 254) // 255


def count_pads(cols, vals):
    """k_count_pads: gaps are measured between the entries whose stored value is not zero."""
    n = 0
    idx = int(cols[0]) if len(cols) else 0
    for c, v in zip(cols[1:], vals[1:]):
        if v == 0.0:
            continue
        n += bridge_entries(int(c) - idx)
        idx = int(c)
    return n


def lanes_and_chunks(slots, cap):
    """StreamDesc::add_group: lanes per row (power of two) and chunks per lane."""
    nch1 = (slots + 3) // 4
    tpr = 1
    while (nch1 + tpr - 1) // tpr > cap and tpr < 32:
        tpr *= 2
    return tpr, (nch1 + tpr - 1) // tpr


def fill(cols, vals, tpr, nch):
    """k_fill_stream, IW == 1: per lane (start index, [(delta byte, value)] * 4*nch)."""
    out = []
    n = len(cols)
    for part in range(tpr):
        first, last = part * 4 * nch, (part + 1) * 4 * nch
        e = jump_left = idx = base = 0
        entries = []
        for s in range(last):
            v, dbyte = 0.0, 0
            if e < n:
                if jump_left > 0:
                    piece = min(jump_left, 255)
                    dbyte = piece
                    idx += piece << 8
                    jump_left -= piece
                else:
                    c, sv = int(cols[e]), float(vals[e])
                    if e == 0:
                        idx = c
                    if sv != 0.0:
                        dbyte, idx, v = c - idx, c, sv
                    e += 1
                    if e < n and float(vals[e]) != 0.0:
                        jump_left = (int(cols[e]) - idx) >> 8
            if s < first:
                continue
            if s == first:
                base, dbyte = idx, 0
            entries.append((dbyte, v))
        out.append((base, entries))
    return out


def decode(lanes):
    """chunk_load, IW == 1: running index per lane; returns {index: summed value}."""
    acc = {}
    for base, entries in lanes:
        cur = base
        for d, v in entries:
            assert 0 <= d <= 255
            cur += d << 8 if v == 0.0 else d
            if v != 0.0:
                acc[cur] = acc.get(cur, 0.0) + v
    return acc


def test_rows_survive_expansion_blocking_and_decoding():
    rng = np.random.default_rng(0)
    for trial in range(800):
        B = int(rng.choice([300, 5000, 70000, 40_000_000]))
        L = int(rng.integers(0, 60))
        cols = np.sort(rng.choice(B, size=min(L, B), replace=False)) if L else np.array([], dtype=int)
        vals = rng.random(len(cols)) + 0.1
        if len(cols) and trial % 3 == 0:                       # explicit zeros stored in D, also first and last
            vals[rng.random(len(cols)) < 0.3] = 0.0
            if trial % 6 == 0:
                vals[0] = vals[-1] = 0.0
        slots = len(cols) + count_pads(cols, vals)
        tpr, nch = lanes_and_chunks(slots, int(rng.choice([1, 2, 5, 24, 48])))
        assert tpr * nch * 4 >= slots                          # every expanded entry has a place
        got = decode(fill(cols, vals, tpr, nch))
        assert got == {int(c): float(v) for c, v in zip(cols, vals) if v != 0.0}


def test_gap_edges():
    for gap, pads in ((0, 0), (1, 0), (255, 0), (256, 1), (511, 1), (5184, 1), (65535, 1), (65536, 2),
                      (255 * 256 + 255, 1), (2 * 255 * 256 + 255, 2), (2 * 255 * 256 + 256, 3)):
        assert bridge_entries(gap) == pads
        cols = np.array([7, 7 + gap])
        vals = [1.0, 2.0]
        assert count_pads(cols, vals) == pads
        tpr, nch = lanes_and_chunks(2 + pads, 48)
        lanes = fill(cols, vals, tpr, nch)
        want = {7: 3.0} if gap == 0 else {7: 1.0, 7 + gap: 2.0}
        assert decode(lanes) == want
        assert max(d for _, ent in lanes for d, _ in ent) <= 255


def test_bridges_on_a_config_like_row():
    """A voxel seen from 180 control points of 81 beamlets each but hidden from 64 of them in a row: the
    5 184-wide gap costs one entry (the 255-per-entry chain of the first layout needed 20)."""
    cols = np.concatenate([np.arange(0, 30) * 81 + 40, np.arange(94, 180) * 81 + 40])
    vals = np.ones(len(cols))
    assert count_pads(cols, vals) == 1
