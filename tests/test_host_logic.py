"""Host-side logic of the product (no GPU): library symbols, aperture bookkeeping, price-row
descriptors, updateOpenAperture - each against the oracle / golden fixtures."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import vmat_oracle as O
from vmatwpencode_b200 import VMATlibrary as vl
from vmatwpencode_b200 import _lib
from vmatwpencode_b200._lib import PriceRow
from vmatwpencode_b200.synth import make_case
from helpers import load_golden, all_candidates, select, oracle_state_from_golden
from k5_emulator import nodes_from_row

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "vmat_b200.h")).read()
    declared = set(re.findall(r"\b(vmat_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert os.path.isfile(_lib.LIB_PATH), "libvmat_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/vmat_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.load().vmat_version() >= 100


def test_struct_layouts_match_header():
    assert ctypes.sizeof(PriceRow) == 56
    assert ctypes.sizeof(_lib.PriceParams) == 64          # C, C2, C3, b, M, N, variant, reserved[5]
    assert ctypes.sizeof(_lib.Options) == 80


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product must raise, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    case = make_case(V=300, nb=3, M=2, N=4, density=0.1, seed=1, gastep=10)
    with pytest.raises(_lib.VmatError):
        vl.bind(case)
    data = vl.bind(case, gpu=False)
    with pytest.raises(_lib.VmatError):
        data.calcDose()


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    for sub in ("vmatwpencode_b200", "tools"):
        pkg = os.path.join(ROOT, sub)
        for fn in os.listdir(pkg):
            if fn.endswith(".py"):
                src = open(os.path.join(pkg, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), (sub, fn)
    top = open(os.path.join(ROOT, "VMATlibrary.py")).read()
    assert not re.search(r"^\s*(from|import)\s+oracle", top, re.M)


@pytest.mark.parametrize("name", ["evalprice_A", "evalprice_B", "evalprice_C", "evalprice_D"])
def test_update_open_aperture_matches_golden(name):
    z, case = load_golden(name)
    data = vl.bind(case, gpu=False)
    all_candidates(data, case)
    for beam in z["selected"]:
        beam = int(beam)
        oam, dm, stg = select(data, case, beam, z[f"sel{beam}_l"], z[f"sel{beam}_r"], vl.updateOpenAperture)
        assert np.array_equal(oam, z[f"sel{beam}_oam"])
        assert np.array_equal(dm, z[f"sel{beam}_diag"])
        assert np.array_equal(np.asarray(stg, dtype=float), z[f"sel{beam}_strength"])


def test_update_open_aperture_random_vs_oracle():
    case = make_case(V=500, nb=4, M=4, N=9, density=0.08, seed=44, gastep=10, ragged=True)
    data = vl.bind(case, gpu=False)
    st = O.OracleState(case)
    rng = np.random.default_rng(9)
    for trial in range(60):
        beam = int(rng.integers(0, case.nb))
        l = [float(rng.integers(-1, case.N - 1)) + (rng.choice([0.0, 0.25, 0.5, 0.995]) if trial % 2 else 0.0)
             for _ in range(case.M)]
        r = [min(case.N + 1.0, lv + float(rng.integers(0, 5)) + (rng.choice([0.0, 0.3, 0.75, 0.005]) if trial % 3 else 0.0))
             for lv in l]
        for state in (data, st):
            state.llist[beam] = list(l)
            state.rlist[beam] = list(r)
        a = vl.updateOpenAperture(beam)
        b = O.update_open_aperture(st, beam)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and list(a[2]) == list(b[2])


def test_fvalidbeamlets_and_neighbours():
    case = make_case(V=500, nb=6, M=4, N=7, density=0.08, seed=45, gastep=10, ragged=True)
    data = vl.bind(case, gpu=False)
    st = O.OracleState(case)
    for i in range(case.nb):
        for m in range(case.M):
            vb, special = vl.fvalidbeamlets(m, i)
            assert np.array_equal(vb, O.valid_beamlets(st, m, i))
            assert special[0] == vb.min() - 1 and special[-1] == vb.max() + 1
    for state in (data, st):
        all_candidates(state, case)
        for b in (1, 4):
            state.caligraphicC.insertAngle(b, case.angles[b])
            state.notinC.removeIndex(b)
    for i in (0, 2, 3, 5):
        assert vl._neighbours(i) == O.neighbours(st, i)


@pytest.mark.parametrize("name", ["evalprice_A", "evalprice_B"])
def test_price_row_descriptors_enumerate_the_oracle_network(name):
    z, case = load_golden(name)
    st = oracle_state_from_golden(z, case)
    data = vl.bind(case, gpu=False)
    all_candidates(data, case)
    for beam in z["selected"]:
        beam = int(beam)
        select(data, case, beam, z[f"sel{beam}_l"], z[f"sel{beam}_r"], vl.updateOpenAperture)
    for k in range(int(z["n_price"])):
        C, C2, C3, b, vmax, speedlim, bw = [float(v) for v in z[f"price{k}_params"]]
        for idx in z[f"price{k}_idx"]:
            idx = int(idx)
            predec, succ, am, ap = vl._neighbours(idx)
            rows = (PriceRow * case.M)()
            vl._fill_price_rows(rows, 0, am, ap, vmax, speedlim, predec, succ, case.N, case.M, idx, bw)
            lcm, rcm, lcp, rcp, vmaxm, vmaxp = vl._neighbour_leaves(predec, succ, case.N, case.M, vmax)
            win = O.leaf_window(case.N, case.M, lcm, lcp, rcm, rcp, vmaxm * (am / speedlim) / bw,
                                vmaxp * (ap / speedlim) / bw, am, ap)
            offset = 0
            for m in range(case.M):
                got = nodes_from_row(rows[m])
                want = [(float(l), [float(r) for r in rs]) for l, rs in win[m]]
                assert got == want, (name, k, idx, m)
                vb = O.valid_beamlets(st, m, idx)
                assert rows[m].vb_min == vb.min() and rows[m].vb_mask == sum(1 << int(v) for v in vb)
                assert rows[m].grad_offset == (0 if m == 0 else offset)
                offset = len(vb) - 1 if m == 0 else offset + len(vb)


def test_kernel_algorithm_twin_matches_oracle_on_random_states():
    """The pricing kernel's algorithm (tests/k5_emulator.py, fed by the product's descriptors)
    against the oracle, on wide ragged grids with loose and tight leaf speeds."""
    from k5_emulator import emulate_price
    for seed, (nb, M, N, ragged, gastep) in enumerate([(10, 6, 12, True, 2), (8, 5, 9, False, 10), (9, 4, 20, True, 1)]):
        case = make_case(V=1500, nb=nb, M=M, N=N, density=0.05, seed=80 + seed, gastep=gastep, ragged=ragged)
        st = O.OracleState(case)
        data = vl.bind(case, gpu=False)
        rng = np.random.default_rng(seed)
        for state in (st, data):
            all_candidates(state, case)
        y = np.zeros(nb)
        for bm in sorted(rng.choice(nb, size=3, replace=False).tolist()):
            l = [int(rng.integers(-1, N - 2)) for _ in range(M)]
            r = [int(min(N, lv + rng.integers(1, 6))) for lv in l]
            select(st, case, bm, l, r, lambda k: O.update_open_aperture(st, k))
            select(data, case, bm, l, r, vl.updateOpenAperture)
            y[bm] = rng.random() * 2
        O.calc_obj_grad_literal(st, y.copy())
        for C, vmax in [(0.0, 2.25), (0.03, 2.25), (1.0, 2.25), (0.03, 0.2), (0.2, 0.02)]:
            for i0 in st.notinC.loc:
                p0, l0, r0, _ = O.price_candidate(st, i0, C, 1.0, 1.0, 0.5, vmax, 0.83, N, M, 0.5)
                predec, succ, am, ap = vl._neighbours(i0)
                rows = (PriceRow * M)()
                vl._fill_price_rows(rows, 0, am, ap, vmax, 0.83, predec, succ, N, M, i0, 0.5)
                bg = st.Dlist[i0] @ st.voxelgradient
                p, l, r = emulate_price(rows, M, C, 1.0, 1.0, 0.5, bg)
                assert (l, r) == (l0, r0) and p == p0, (seed, C, vmax, i0)


def test_kelly_measure_vs_oracle():
    rng = np.random.default_rng(12)
    for _ in range(200):
        m = int(rng.integers(1, 9))
        l = [int(rng.integers(-1, 10)) for _ in range(m)]
        r = [lv + int(rng.integers(1, 8)) for lv in l]
        assert abs(vl.kelly_measure(l, r) - O.kelly_measure(l, r)) <= 1e-14 * abs(O.kelly_measure(l, r))


def test_quad_helpers_vs_oracle():
    rng = np.random.default_rng(13)
    nv, ns = 500, 6
    owner = rng.integers(0, ns, size=nv)
    regions = [vl.region(s, np.flatnonzero(owner == s), np.flatnonzero(owner == s), s < 2) for s in range(ns)]
    regions[3] = vl.region(3, np.array([], dtype=int), np.array([], dtype=int), False)     # an empty structure
    fd = rng.random(3 * ns) * 10
    prio = rng.permutation(ns).tolist()
    a = vl.quadHelpers(fd, regions, nv, prio)
    b = O.quad_helpers(fd, regions, nv, prio)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_subsample_case_keeps_rows_and_rescales_weights():
    from vmatwpencode_b200.synth import subsample_case, voxel_coordinates
    case = make_case(V=900, nb=4, M=3, N=4, density=0.05, seed=12)
    xyz = voxel_coordinates(case.V, case.M)
    assert xyz.shape == (900, 3) and len({tuple(p) for p in xyz}) == 900 and np.array_equal(xyz, np.round(xyz))
    same = subsample_case(case, np.arange(case.V)[::-1])          # every voxel, any order given
    assert np.array_equal(same.over, case.over) and np.array_equal(same.under, case.under)
    assert (same.D != case.D.tocsr()).nnz == 0
    rows = np.arange(0, case.V, 3)
    sub = subsample_case(case, rows)
    assert sub.V == len(rows) and np.array_equal(sub.thr, case.thr[rows])
    D = case.D.tocsr()
    assert (sub.D != D[rows]).nnz == 0
    for s in np.unique(case.struct_id):
        full, kept = (case.struct_id == s).sum(), (sub.struct_id == s).sum()
        if kept:
            sel = sub.struct_id == s
            assert np.allclose(sub.over[sel], case.over[rows][sel] * full / kept, rtol=1e-15)


def test_lean_lbfgsb_loop_equals_scipy_minimize_bitwise():
    """VMATlibrary._lbfgsb_lean drives scipy's compiled L-BFGS-B routine with the loop, work arrays and stopping rules
    of scipy.optimize.minimize(method="L-BFGS-B"): same iterates, objective, iteration and evaluation counts, bit for
    bit (the greedy loop's aperture sequences depend on it)."""
    import warnings
    from scipy.optimize import minimize
    assert vl._lean_available()
    rng = np.random.default_rng(3)
    for n, ftol, maxiter in ((40, 1e-1, 200), (25, 1e-9, 7), (60, 1e-6, 1000)):
        A = rng.random((n + 20, n))
        b = rng.random(n + 20) * 5

        def fg(x):
            r = A @ x - b
            rp = np.maximum(r, 0)
            return float((rp ** 3).sum() + (r ** 2).sum()), A.T @ (3 * rp ** 2 + 2 * r)

        up = np.where(rng.random(n) > 0.3, 12.048, 0.0)
        lo = np.zeros(n)
        x0 = rng.random(n) * up
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = minimize(fg, x0, method="L-BFGS-B", jac=True, bounds=list(zip(lo, up)),
                           options={"ftol": ftol, "maxiter": maxiter})
        got = vl._lbfgsb_lean(fg, x0, lo, up, ftol=ftol, maxiter=maxiter)
        assert np.array_equal(ref.x, got.x) and ref.fun == got.fun
        assert (ref.nit, ref.nfev, ref.status) == (got.nit, got.nfev, got.status)
