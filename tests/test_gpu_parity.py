"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the
C ABI (ctypes), against the oracle and against the fixtures frozen from the reference.

Tolerances: leaf positions and candidate choice bit-exact; dose / objective / gradients
within 1e-5 relative (max-norm) for fp32 accumulation (BASELINE.json north_star), 1e-11 for
fp64 accumulation."""
import os

import numpy as np
import pytest

from oracle import vmat_oracle as O
from vmatwpencode_b200 import VMATlibrary as vl
from vmatwpencode_b200 import _lib
from vmatwpencode_b200.plan import DosePlan
from vmatwpencode_b200.synth import make_case, make_config
from helpers import load_golden, all_candidates, select, oracle_state_from_golden, relerr

pytestmark = pytest.mark.gpu

TOL = {"f32": 1e-5, "f64": 1e-11}


def product_state_from_golden(z, case, **kw):
    data = vl.bind(case, **kw)
    all_candidates(data, case)
    for beam in z["selected"]:
        beam = int(beam)
        select(data, case, beam, z[f"sel{beam}_l"], z[f"sel{beam}_r"], vl.updateOpenAperture)
    return data


def test_native_library_is_the_in_tree_build():
    lib = _lib.load()
    assert os.path.dirname(_lib.LIB_PATH).endswith("vmatwpencode_b200")
    assert lib.vmat_version() >= 100


@pytest.mark.parametrize("acc", ["f64", "f32"])
@pytest.mark.parametrize("name", ["evalprice_A", "evalprice_B", "evalprice_C", "evalprice_D"])
def test_eval_against_reference_fixture(name, acc):
    z, case = load_golden(name)
    data = product_state_from_golden(z, case, store="f32", acc=acc)
    f, g = vl.calcObjGrad(np.array(z["y"]))
    tol = TOL[acc]
    assert abs(f - float(z["f"])) <= tol * abs(float(z["f"]))
    assert g.shape == (case.nb, 1)
    assert relerr(g, z["g"]) <= tol
    assert relerr(data.currentDose, z["dose"]) <= tol
    assert relerr(data.voxelgradient, z["vg"]) <= tol
    data.plan.close()


@pytest.mark.parametrize("name", ["evalprice_A", "evalprice_B", "evalprice_C", "evalprice_D"])
def test_pricing_against_reference_fixture(name):
    z, case = load_golden(name)
    data = product_state_from_golden(z, case, store="f32", acc="f64")
    vl.calcObjGrad(np.array(z["y"]))
    for k in range(int(z["n_price"])):
        C, C2, C3, b, vmax, speedlim, bw = [float(v) for v in z[f"price{k}_params"]]
        # one candidate at a time through the reference's per-candidate entry point ...
        for n, idx in enumerate(z[f"price{k}_idx"]):
            p, l, r, i = vl.parallelizationPricingProblem(int(idx), C, C2, C3, b, vmax, speedlim, case.N, case.M, bw)
            assert [int(v) for v in l] == list(z[f"price{k}_l"][n]), (name, k, idx)
            assert [int(v) for v in r] == list(z[f"price{k}_r"][n]), (name, k, idx)
            assert abs(p - z[f"price{k}_p"][n]) <= 1e-11 * max(1.0, abs(z[f"price{k}_p"][n]))
        # ... and all candidates in one launch
        pstar, l, r, best = vl.PricingProblem(C, C2, C3, b, vmax, speedlim, case.N, case.M, bw)
        want = int(np.argmin(z[f"price{k}_p"]))
        assert best == int(z[f"price{k}_idx"][want])
        assert [int(v) for v in l] == list(z[f"price{k}_l"][want])
        assert [int(v) for v in r] == list(z[f"price{k}_r"][want])
    data.plan.close()


@pytest.mark.parametrize("name", ["colgen_A", "colgen_B", "colgen_C", "colgen_E"])
def test_greedy_sequence_against_reference_fixture(name):
    z, case = load_golden(name)
    data = vl.bind(case, store="f32", acc="f64")
    vl.eliminationPhase = bool(z["elimination"])
    try:
        vl.colGen(float(z["C"]), False, int(z["kappasize"]), max_iter=int(z["max_iter"]))
    finally:
        vl.eliminationPhase = False
    hist = data.history
    assert [h["idx"] for h in hist] == list(z["idx"])
    assert [k for h in hist for k in h["removed"]] == list(z["removed_flat"])
    if "kelly" in z.files:
        got = {k: v for k, v in enumerate(vl.penalizationweights) if v is not None}
        assert sorted(got) == list(z["kelly_idx"])
        assert np.allclose([got[k] for k in sorted(got)], z["kelly"], rtol=1e-14, atol=0)
    for h, l, r, p, x, fun in zip(hist, z["l"], z["r"], z["pstar"], z["x"], z["fun"]):
        assert h["l"] == list(l) and h["r"] == list(r)          # bit-exact leaf positions
        assert abs(h["pstar"] - p) <= 1e-9 * abs(p)
        assert abs(h["fun"] - fun) <= 1e-9 * abs(fun)
        assert np.abs(h["x"] - x).max() <= 1e-7 * max(1.0, np.abs(x).max())
    data.plan.close()


def random_weights(case, rng, frac_selected=0.5):
    w_dose = np.zeros(case.B)
    w_grad = np.zeros(case.B)
    y = np.zeros(case.nb)
    for i in range(case.nb):
        if rng.random() < frac_selected:
            lo, hi = int(case.beam_ptr[i]), int(case.beam_ptr[i + 1])
            w_dose[lo:hi] = (rng.random(hi - lo) > 0.4) * np.round(rng.random(hi - lo) * 4) / 4
            w_grad[lo:hi] = (rng.random(hi - lo) > 0.4) * 1.0
            y[i] = rng.random() * 3
    return y, w_dose, w_grad


@pytest.mark.parametrize("store,acc,val_dtype,tol", [
    ("f32", "f32", "f32", 1e-5), ("f32", "f64", "f32", 1e-11), ("f64", "f64", "f64", 1e-11),
    ("bf16", "f32", "bf16", 1e-5), ("bf16", "f64", "bf16", 1e-11), ("f64", "f32", "f32", 1e-5)])
@pytest.mark.parametrize("shape", [
    dict(V=50_000, nb=24, M=6, N=8, density=0.01, seed=61),
    dict(V=4099, nb=7, M=3, N=5, density=0.05, seed=62, ragged=True),       # V % 32 != 0, ragged rows
    dict(V=33, nb=2, M=1, N=3, density=0.3, seed=63),                       # tiny
    dict(V=20_000, nb=10, M=4, N=6, density=0.002, seed=64, model="uniform"),   # many empty voxel rows
])
@pytest.mark.parametrize("fused", [2, 1])          # 2: the single fused evaluation kernel, 1: the three kernels K1/K2/K3
def test_eval_random_state_vs_oracle(shape, store, acc, val_dtype, tol, fused):
    case = make_case(val_dtype=val_dtype, **shape)
    st = O.OracleState(case)
    rng = np.random.default_rng(shape["seed"])
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, store=store, acc=acc,
                    tile_voxels=4096 if shape["V"] > 10_000 else 64, fused=fused)
    for trial in range(3):
        y, w_dose, w_grad = random_weights(case, rng, frac_selected=[0.5, 1.0, 0.0][trial])
        plan.set_weights(w_dose, w_grad)
        f, g = plan.eval(y)
        f0, g0, d0, vg0, gb0 = O.calc_obj_grad_clean(st, y, w_dose, w_grad)
        assert abs(f - f0) <= tol * abs(f0)
        assert relerr(g, g0) <= tol
        assert relerr(plan.dose(), d0) <= tol
        assert relerr(plan.voxelgrad(), vg0) <= tol
        assert relerr(plan.beamlet_grad(), gb0) <= tol
    plan.close()


@pytest.mark.parametrize("opts", [dict(use_graph=False), dict(tile_voxels=128), dict(tile_voxels=64, k2_tasks=7),
                                  dict(k1_threads=128), dict(k2_threads=128, k2_tasks=1000), dict(pdl=False), dict(host_path=3), dict(host_path=4), dict(host_path=6),
                                  dict(chunk_cap=2), dict(chunk_cap=1000), dict(index_bytes=4), dict(tile_voxels=64, chunk_cap=1),
                                  dict(k2_chunk_cap=1, k2_tasks=5), dict(k2_chunk_cap=3, tile_voxels=128), dict(claim_lead=1), dict(claim_lead=100),
                                  dict(k2_threads=256, k2_stages=6), dict(k2_threads=512, k2_stages=4, tile_voxels=256),
                                  dict(index_bytes=1), dict(index_bytes=1, chunk_cap=2, k2_chunk_cap=3), dict(index_bytes=1, tile_voxels=64, k2_tasks=7),
                                  dict(index_bytes=1, k1_threads=128, k2_threads=256),
                                  # the fused evaluation kernel: grid sizes (fewer CTAs than control points, one CTA), short
                                  # rings, one-chunk items, tiny tiles (many groups, tile buffers change with every item)
                                  dict(fused=2), dict(fused=2, k2_tasks=3), dict(fused=2, k2_tasks=1, k1_stages=2),
                                  dict(fused=2, k1_stages=3, tile_voxels=64, k2_tasks=5), dict(fused=2, chunk_cap=1, k2_chunk_cap=1),
                                  dict(fused=2, chunk_cap=1000, k2_chunk_cap=1000, tile_voxels=128), dict(fused=2, index_bytes=4, claim_lead=1),
                                  dict(fused=2, pdl=False, use_graph=False), dict(fused=2, tile_voxels=64, claim_lead=100, k2_tasks=200),
                                  dict(fused=1), dict(fused=1, k2_tasks=7, tile_voxels=64),
                                  # the transposed kernel's dynamic tail: none, half of the work, with few CTAs / tiny tiles
                                  dict(k2_pool_pct=-1), dict(k2_pool_pct=50, tile_voxels=64), dict(k2_pool_pct=30, k2_tasks=7),
                                  dict(k2_pool_pct=90, k2_chunk_cap=2, tile_voxels=128)])
def test_plan_options_do_not_change_results(opts):
    case = make_case(V=9000, nb=8, M=4, N=6, density=0.02, seed=71)
    st = O.OracleState(case)
    y, w_dose, w_grad = random_weights(case, np.random.default_rng(2), 0.7)
    f0, g0, d0, vg0, gb0 = O.calc_obj_grad_clean(st, y, w_dose, w_grad)
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc="f64", **opts)
    plan.set_weights(w_dose, w_grad)
    f, g = plan.eval(y)
    assert abs(f - f0) <= 1e-11 * abs(f0) and relerr(g, g0) <= 1e-11 and relerr(plan.beamlet_grad(), gb0) <= 1e-11
    # resident path: same numbers without host round trips
    plan.set_intensities(y)
    plan.eval_async()
    f2, g2 = plan.fetch_result()
    assert f2 == f and np.array_equal(g2, g)
    plan.close()


@pytest.mark.parametrize("store,acc,tol", [("f32", "f64", 1e-11), ("f32", "f32", 1e-5), ("f64", "f64", 1e-11), ("bf16", "f32", 1e-5)])
def test_byte_delta_indices_bridge_large_gaps(store, acc, tol):
    """index_bytes=1 stores column indices as byte deltas.  A matrix far too sparse for that (70 000
    beamlets, ~30 nonzeros per voxel row: gaps in the thousands, each bridged by a zero-valued entry
    whose delta counts units of 256; also more beamlets than 16 bits address) still evaluates to the
    oracle's numbers."""
    import scipy.sparse as sp
    rng = np.random.default_rng(21)
    V, nb, per = 3000, 10, 7000
    B = nb * per
    nnz = 90_000
    rows, cols = rng.integers(0, V, nnz), rng.integers(0, B, nnz)
    rows[:200] = 5; cols[:200] = np.arange(200) * 2 + 11            # one long dense-ish row: no bridges
    cols[200:260] = rng.choice([0, 255, 256, 510, 511, 512, 69_999], 60); rows[200:260] = np.arange(60) % 3 + 7   # exact gap edges
    vals = np.float32(0.05 + rng.random(nnz)).astype(np.float64)
    if store == "bf16":
        import torch
        vals = torch.tensor(vals, dtype=torch.float32).to(torch.bfloat16).to(torch.float64).numpy()
    D = sp.csr_matrix((vals, (rows, cols)), shape=(V, B))
    D.sum_duplicates()
    if store == "bf16":
        import torch
        D.data = torch.tensor(D.data, dtype=torch.float32).to(torch.bfloat16).to(torch.float64).numpy()
    else:
        D.data = np.float32(D.data).astype(np.float64)
    # stored zeros (the layout marks bridge entries by a zero value, so real entries that are zero - or become zero
    # in the storage type - must not move the running index): explicit zeros, also first and last of a row, and
    # values that underflow to zero in fp32 / bf16 storage
    zero_at = rng.choice(D.nnz, 3000, replace=False)
    D.data[zero_at[:2000]] = 0.0
    D.data[zero_at[2000:2500]] = 1e-60
    D.data[zero_at[2500:]] = 1e-45       # a float32 denormal: a real entry in fp32 storage, zero in bf16 storage
    D.data[D.indptr[7]] = 0.0; D.data[D.indptr[8] - 1] = 0.0; D.data[D.indptr[9]:D.indptr[9] + 3] = 0.0
    beam_ptr = np.arange(nb + 1, dtype=np.int32) * per
    thr = rng.random(V); over = rng.random(V) / V; under = rng.random(V) / V
    y = rng.random(nb) * 2
    w_dose = (rng.random(B) > 0.5) * rng.random(B); w_grad = (rng.random(B) > 0.5) * 1.0
    x = np.repeat(y, per) * w_dose
    d0 = D @ x
    o, u = np.maximum(d0 - thr, 0), np.maximum(thr - d0, 0)
    f0 = float((over * o * o + under * u * u).sum())
    vg0 = 4 * (over * o - under * u)
    gb0 = D.T @ vg0
    g0 = (gb0 * w_grad).reshape(nb, per).sum(axis=1)
    plan = DosePlan(D, beam_ptr, thr, over, under, store=store, acc=acc, index_bytes=1)
    assert plan.info()["index_bytes"] == 1
    plan.set_weights(w_dose, w_grad)
    f, g = plan.eval(y)
    assert abs(f - f0) <= tol * abs(f0) and relerr(g, g0) <= tol
    assert relerr(plan.dose(), d0) <= tol and relerr(plan.beamlet_grad(), gb0) <= tol
    assert plan.guard_status() is None          # no running index ever left the gathered vector
    plan.close()


def test_default_index_width_follows_size_precision_and_sortedness():
    """index_bytes=0: fp32 storage with fp32 accumulation stores byte deltas from 5 M nonzeros on; 8-byte
    accumulation keeps absolute indices; so does a caller whose voxel-major rows are not sorted (byte deltas
    need sorted rows, and choosing them was the library's idea) - with the same results."""
    import torch
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    V, nb, per, L = 52_000, 20, 110, 100
    B = nb * per
    first = rng.integers(0, B - 4 * L, V)
    step = rng.integers(1, 5, V)
    cols = first[:, None] + step[:, None] * np.arange(L)[None, :]                 # sorted, distinct within a row
    vals = np.float32(0.05 + rng.random((V, L)))
    indptr = np.arange(V + 1, dtype=np.int64) * L
    D = sp.csr_matrix((vals.ravel().astype(np.float64), cols.ravel(), indptr), shape=(V, B))
    assert D.nnz >= 5_000_000
    Dc = D.tocsc(); Dc.sort_indices()
    beam_ptr = np.arange(nb + 1, dtype=np.int32) * per
    thr = rng.random(V) * 5; over = rng.random(V) / V; under = rng.random(V) / V
    y = rng.random(nb)
    w = np.ones(B)
    dev = lambda a, t: torch.as_tensor(np.ascontiguousarray(a), dtype=t, device="cuda")
    def parts(c, v):
        return dict(vrowptr=dev(indptr, torch.int64), vcol=dev(c.ravel(), torch.int32), vval=dev(v.ravel(), torch.float32),
                    browptr=dev(Dc.indptr, torch.int64), bcol=dev(Dc.indices, torch.int32), bval=dev(Dc.data, torch.float32))
    out = {}
    for name, c, v, acc in (("sorted", cols, vals, "f32"), ("reversed", cols[:, ::-1], vals[:, ::-1], "f32"), ("f64", cols, vals, "f64")):
        plan = DosePlan(parts(c, v), beam_ptr, thr, over, under, store="f32", acc=acc)
        plan.set_weights(w, w)
        f, g = plan.eval(y)
        out[name] = (plan.info()["index_bytes"], f, g.copy(), plan.dose())
        plan.close()
    assert out["sorted"][0] == 1 and out["reversed"][0] == 2 and out["f64"][0] == 2
    for name in ("sorted", "reversed"):
        assert abs(out[name][1] - out["f64"][1]) <= 1e-5 * abs(out["f64"][1])
        assert relerr(out[name][2], out["f64"][2]) <= 1e-5 and relerr(out[name][3], out["f64"][3]) <= 1e-5
    with pytest.raises(_lib.VmatError, match="sorted"):                           # asked for explicitly: an error
        DosePlan(parts(cols[:, ::-1], vals[:, ::-1]), beam_ptr, thr, over, under, store="f32", acc="f32", index_bytes=1)


def test_evaluation_is_bitwise_reproducible_and_apertures_update():
    case = make_case(V=30_000, nb=12, M=5, N=8, density=0.01, seed=72)
    rng = np.random.default_rng(3)
    y, w_dose, w_grad = random_weights(case, rng, 0.6)
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc="f32")
    plan.set_weights(w_dose, w_grad)
    f1, g1 = plan.eval(y)
    gb1 = plan.beamlet_grad()
    for _ in range(3):
        f2, g2 = plan.eval(y)
        assert f2 == f1 and np.array_equal(g1, g2) and np.array_equal(gb1, plan.beamlet_grad())
    # closing every control point gives the zero-dose objective sum(under * thr^2)
    for i in range(case.nb):
        plan.set_aperture(i, None, None)
    f0, g0 = plan.eval(y)
    want = float(np.sum(case.under * case.thr * case.thr))
    assert abs(f0 - want) <= 1e-5 * want and np.all(g0 == 0)
    assert np.all(plan.dose() == 0)
    plan.close()


def test_argument_errors_are_reported_not_ignored():
    case = make_case(V=500, nb=3, M=2, N=4, density=0.1, seed=73)
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under)
    with pytest.raises(_lib.VmatError):
        plan.set_aperture(99, None, None)
    with pytest.raises(_lib.VmatError):
        plan.set_aperture(0, np.zeros(3), np.zeros(3))
    with pytest.raises(_lib.VmatError):
        DosePlan(case.D, case.beam_ptr[:-1], case.thr, case.over, case.under)
    plan.close()


def test_malformed_csr_is_rejected_at_plan_build():
    """Out-of-range or unsorted indices would become wild shared-memory reads; the plan refuses them."""
    import torch
    case = make_case(V=600, nb=3, M=2, N=4, density=0.1, seed=74)
    Dr = case.D.tocsr(); Dr.sort_indices()
    Dc = Dr.tocsc(); Dc.sort_indices()
    dev = lambda a, t: torch.as_tensor(np.ascontiguousarray(a), dtype=t, device="cuda")
    def parts():
        return dict(vrowptr=dev(Dr.indptr, torch.int64), vcol=dev(Dr.indices, torch.int32), vval=dev(Dr.data, torch.float32),
                    browptr=dev(Dc.indptr, torch.int64), bcol=dev(Dc.indices, torch.int32), bval=dev(Dc.data, torch.float32))
    args = (case.beam_ptr, case.thr, case.over, case.under)
    DosePlan(parts(), *args).close()                       # the untouched payload is fine
    bad = parts(); bad["vcol"][5] = case.B                 # beamlet index out of range
    with pytest.raises(_lib.VmatError, match="out of range"):
        DosePlan(bad, *args)
    bad = parts(); bad["bcol"][3] = -1
    with pytest.raises(_lib.VmatError, match="out of range"):
        DosePlan(bad, *args)
    bad = parts()
    row = int(np.argmax(np.diff(Dc.indptr) >= 2)); k = int(Dc.indptr[row])
    bad["bcol"][k], bad["bcol"][k + 1] = bad["bcol"][k + 1].clone(), bad["bcol"][k].clone()
    with pytest.raises(_lib.VmatError, match="sorted"):
        DosePlan(bad, *args)
    bad = parts(); bad["vrowptr"][4] = bad["vrowptr"][5] + 1
    with pytest.raises(_lib.VmatError, match="decrease"):
        DosePlan(bad, *args)


def test_pricing_random_states_vs_oracle():
    """Bigger leaf grids, ragged rows, tight leaf speeds (fractional fallbacks), C in {0, >0}."""
    for seed, (nb, M, N, ragged, gastep) in enumerate([(10, 6, 12, True, 2), (8, 9, 9, False, 10), (14, 4, 20, True, 1)]):
        case = make_case(V=6000, nb=nb, M=M, N=N, density=0.03, seed=80 + seed, gastep=gastep, ragged=ragged)
        st = O.OracleState(case)
        data = vl.bind(case, store="f32", acc="f64")
        rng = np.random.default_rng(seed)
        for state in (st, data):
            all_candidates(state, case)
        chosen = sorted(rng.choice(nb, size=3, replace=False).tolist())
        y = np.zeros(nb)
        for bm in chosen:
            l = [int(rng.integers(-1, N - 2)) for _ in range(M)]
            r = [int(min(N, lv + rng.integers(1, 6))) for lv in l]
            select(st, case, bm, l, r, lambda k: O.update_open_aperture(st, k))
            select(data, case, bm, l, r, vl.updateOpenAperture)
            y[bm] = rng.random() * 2
        O.calc_obj_grad_literal(st, y.copy())
        vl.calcObjGrad(y.copy())
        for C, vmax in [(0.0, 2.25), (0.03, 2.25), (1.0, 2.25), (0.03, 0.2), (0.0, 0.05), (0.2, 0.02)]:
            ref = [O.price_candidate(st, i, C, 1.0, 1.0, 0.5, vmax, 0.83, N, M, 0.5) for i in st.notinC.loc]
            for (p0, l0, r0, i0) in ref:
                p, l, r, i = vl.parallelizationPricingProblem(i0, C, 1.0, 1.0, 0.5, vmax, 0.83, N, M, 0.5)
                assert [int(v) for v in l] == l0 and [int(v) for v in r] == r0, (seed, C, vmax, i0)
                assert abs(p - p0) <= 1e-10 * max(1.0, abs(p0))
            got = vl.PricingProblem(C, 1.0, 1.0, 0.5, vmax, 0.83, N, M, 0.5)
            want = O.pricing_problem(st, C, 1.0, 1.0, 0.5, vmax, 0.83, N, M, 0.5)
            assert got[3] == want[3] and [int(v) for v in got[1]] == want[1] and [int(v) for v in got[2]] == want[2]
        data.plan.close()


def test_pricing_before_any_evaluation_is_an_error():
    case = make_case(V=500, nb=3, M=2, N=4, density=0.1, seed=74)
    data = vl.bind(case)
    all_candidates(data, case)
    with pytest.raises(_lib.VmatError):
        vl.PricingProblem(0.0, 1.0, 1.0, 0.5, 2.25, 0.83, case.N, case.M, 0.5)
    data.plan.close()


def test_greedy_loop_config1_to_completion_vs_oracle():
    """BASELINE.json configs[0] (20k voxels x 2 016 beamlets, 36 control points): the whole greedy
    run, until every control point has an aperture or pricing finds nothing: same aperture
    sequence, leaf for leaf, as the oracle."""
    case = make_config("cfg1")
    want = O.col_gen(O.OracleState(case), C=0.01, kappasize=16, max_iter=40, clean=True)
    data = vl.bind(case, store="f32", acc="f64")
    vl.colGen(0.01, False, 16, max_iter=40)
    assert len(want) >= 30
    assert [h["idx"] for h in data.history] == [h["idx"] for h in want]
    for a, b in zip(data.history, want):
        assert a["l"] == b["l"] and a["r"] == b["r"]
        assert abs(a["fun"] - b["fun"]) <= 1e-6 * abs(b["fun"])
    data.plan.close()


def test_greedy_sequence_config2_vs_oracle():
    """BASELINE.json configs[1] at full size (300k voxels x 10 080 beamlets, 180 control points): the
    first greedy iterations pick the same control points with the same leaves as the CPU oracle."""
    case = make_config("cfg2")
    data = vl.bind(case, store="f32", acc="f64")
    vl.colGen(0.01, False, 16, max_iter=3)
    got = [(h["idx"], tuple(h["l"]), tuple(h["r"])) for h in data.history]
    hist = O.col_gen(O.OracleState(case), C=0.01, kappasize=16, max_iter=3, clean=True)
    assert got == [(h["idx"], tuple(h["l"]), tuple(h["r"])) for h in hist]
    f, _ = data.plan.eval(np.asarray(data.currentIntensities, dtype=float))
    data.plan.close()
    with pytest.raises(_lib.VmatError):
        data.plan.eval(np.zeros(case.nb))            # a closed plan reports, it does not crash


def test_kmedoids_known_answer():
    from vmatwpencode_b200 import kmedoids as km
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmedoids.npz"))
    for n in (23, 40, 31):
        start = [int(v) for v in z[f"start_n{n}"]]
        assert km.givemewholelist(start, range(0, n)) == list(z[f"order_n{n}"])
        assert km.distancemin(start, range(0, n)) == float(z[f"dmin_n{n}"])
    # the one known answer the reference embeds (greedyVMAT.py:28)
    assert km.givemewholelist(list(range(6, 178, 11)), range(0, 178)) == list(z["kappa178"])
    # 1-D lists that are not 0..n-1: the reference's candidates are the VALUES of ba, in ba's order
    zp = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmedoids_perm.npz"))
    for n in (19, 26):
        ba = [int(v) for v in zp[f"ba_n{n}"]]
        assert km.givemewholelist([int(v) for v in zp[f"start_n{n}"]], ba) == list(zp[f"order_n{n}"])
    with pytest.raises(IndexError):
        km.givemewholelist([0], [0.5, 1.5, 7.0])
    # coordinates in 3-D (voxel use): against the oracle
    rng = np.random.default_rng(4)
    P = rng.random((300, 3))
    got = km.insertnewelement([0, 5], P)
    assert got == O.km_insertnewelement([0, 5], P)


@pytest.mark.parametrize("acc", ["f32", "f64"])
def test_full_size_config2_properties(acc):
    """BASELINE.json configs[1] at full size (300k voxels x 10 080 beamlets, ~30 M nnz):
    size-independent properties plus a direct comparison with the C oracle."""
    from oracle.c_oracle import COracle
    case = make_config("cfg2")
    rng = np.random.default_rng(7)
    y, w_dose, w_grad = random_weights(case, rng, 0.5)
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc=acc)
    plan.set_weights(w_dose, w_grad)
    tol = TOL[acc]
    f, g = plan.eval(y)
    d, vg, gb = plan.dose(), plan.voxelgrad(), plan.beamlet_grad()
    co = COracle(case)
    f0, g0 = co.eval(y, w_dose, w_grad)
    assert abs(f - f0) <= tol * abs(f0)
    assert relerr(g, g0) <= tol and relerr(d, co.d) <= tol and relerr(vg, co.vg) <= tol and relerr(gb, co.gb) <= tol
    # adjoint identity <D x, vg> == <x, D^T vg>
    x = np.repeat(y, np.diff(case.beam_ptr)) * w_dose
    lhs, rhs = float(np.dot(d, vg)), float(np.dot(x, gb))
    assert abs(lhs - rhs) <= 10 * tol * max(abs(lhs), abs(rhs))
    # dose is linear in y
    plan.eval(2 * y)
    assert relerr(plan.dose(), 2 * d) <= tol
    # aperture gradient is the w_grad-weighted beamlet gradient
    assert relerr(g, np.add.reduceat(w_grad * gb, case.beam_ptr[:-1].astype(np.int64))) <= tol
    plan.close()


@pytest.mark.parametrize("acc", ["f32", "f64"])
def test_full_size_config3_vs_c_oracle(acc):
    """BASELINE.json configs[2]'s matrix at full size (1M voxels x 14 580 beamlets, ~142 M nnz, generated
    on the device): every output against the C oracle run on the SAME arrays copied back to the host."""
    import torch
    from oracle.c_oracle import COracle
    from vmatwpencode_b200.synth import make_device_config
    case = make_device_config("cfg3", device="cuda:0")
    rng = np.random.default_rng(17)
    y, w_dose, w_grad = random_weights(case, rng, 0.5)
    plan = DosePlan(case.csr, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc=acc)
    plan.set_weights(w_dose, w_grad)
    f, g = plan.eval(y)
    d, vg, gb = plan.dose(), plan.voxelgrad(), plan.beamlet_grad()
    plan.close()
    h = {k: v.cpu().numpy() for k, v in case.csr.items()}
    case.csr = None
    torch.cuda.empty_cache()
    co = COracle.from_arrays(case.V, case.B, case.nb, case.beam_ptr, h["vrowptr"], h["vcol"], h["vval"],
                             h["browptr"], h["bcol"], h["bval"], case.thr, case.over, case.under)
    f0, g0 = co.eval(y, w_dose, w_grad)
    tol = TOL[acc]
    assert abs(f - f0) <= tol * abs(f0)
    assert relerr(g, g0) <= tol and relerr(d, co.d) <= tol and relerr(vg, co.vg) <= tol and relerr(gb, co.gb) <= tol


def test_full_size_config4_sampled_check():
    """BASELINE.json configs[3] (4M voxels x 29 160 beamlets, ~1.14 G nnz, device-generated): too large for a
    host oracle in test time, so 10 000 random voxel rows of the dose / voxel gradient and 1 000 random
    beamlets of the beamlet gradient are recomputed in float64 from the device CSR with plain torch index
    arithmetic (bench.sampled_device_check - not the plan's layouts or kernels), the objective from the
    device dose, the aperture gradient from the device beamlet gradient."""
    import torch
    import bench
    from vmatwpencode_b200.synth import make_device_config
    case = make_device_config("cfg4", device="cuda:0")
    y, w_dose, w_grad = bench.bench_weights(case.nb, case.M, case.N, case.beam_ptr, case.B)
    plan = DosePlan(case.csr, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc="f32")
    plan.set_weights(w_dose, w_grad)
    f, g = plan.eval(y)
    chk = bench.sampled_device_check(plan, case.csr, case.thr, case.over, case.under, y, w_dose, w_grad, case.beam_ptr,
                                     f, g, 0, 1, None, torch)
    plan.close()
    case.csr = None
    torch.cuda.empty_cache()
    for k in ("rel_err_f", "rel_err_g", "rel_err_gb", "rel_err_d", "rel_err_vg"):
        assert chk[k] <= 1e-5, (k, chk[k])


def test_back_to_back_evaluations_do_not_race():
    """Evaluations enqueued back to back overlap through programmatic dependent launch (the next
    kernel's prologue runs while the previous kernel drains).  Alternate two intensity vectors
    without any host synchronisation and check every result against the single-shot values."""
    import torch
    case = make_case(V=60_000, nb=24, M=6, N=8, density=0.01, seed=75)
    rng = np.random.default_rng(5)
    y0, w_dose, w_grad = random_weights(case, rng, 0.8)
    ys = [y0, y0 * 0.37 + 0.2, y0 * 1.9]
    for acc in ("f32", "f64"):
        plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc=acc)
        plan.set_weights(w_dose, w_grad)
        want = []
        for y in ys:
            f, g = plan.eval(y)
            want.append(np.concatenate([[f], g]))
        side = torch.cuda.Stream()
        ydev = [torch.tensor(y, dtype=torch.float64, device="cuda") for y in ys]
        ybuf, rbuf = plan.intensities_buffer(), plan.result_buffer()
        n = 90
        out = torch.zeros((n, case.nb + 1), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            for k in range(n):
                ybuf.copy_(ydev[k % 3], non_blocking=True)
                plan.eval_async(side.cuda_stream)
                out[k].copy_(rbuf, non_blocking=True)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        for k in range(n):
            assert np.array_equal(got[k], want[k % 3]), (acc, k)
        # and with nothing between the evaluations at all
        with torch.cuda.stream(side):
            ybuf.copy_(ydev[1], non_blocking=True)
            for k in range(50):
                plan.eval_async(side.cuda_stream)
        torch.cuda.synchronize()
        f, g = plan.fetch_result()
        assert np.array_equal(np.concatenate([[f], g]), want[1])
        plan.close()


def test_device_generated_case_matches_host_generated_case():
    """The on-GPU generator + device-CSR plan construction against the host (scipy) route."""
    import torch
    from vmatwpencode_b200.synth import make_device_case
    kw = dict(V=40_000, nb=12, M=5, N=6, density=0.01, seed=77)
    host = make_case(**kw)
    dev = make_device_case(**kw)
    assert dev.nnz == host.nnz
    D = host.D.tocsr(); D.sort_indices()
    assert np.array_equal(dev.csr["vrowptr"].cpu().numpy(), D.indptr)
    assert np.array_equal(dev.csr["vcol"].cpu().numpy(), D.indices)
    # values agree to the last float32 bit or two (exp/sin/cos differ between CPU and GPU libms)
    assert np.allclose(dev.csr["vval"].cpu().numpy(), D.data, rtol=3e-7, atol=0)
    rng = np.random.default_rng(1)
    y, w_dose, w_grad = random_weights(host, rng, 0.7)
    Dd = __import__("scipy.sparse", fromlist=["csr_matrix"]).csr_matrix(
        (dev.csr["vval"].cpu().numpy().astype(np.float64), D.indices, D.indptr), shape=D.shape)
    host.D = Dd
    # structure membership of a few boundary voxels may differ between CPU and GPU float math:
    # the oracle takes the device case's own penalty tables
    host.thr, host.over, host.under = dev.thr, dev.over, dev.under
    assert np.mean(host.thr != make_case(**kw).thr) < 1e-3
    st = O.OracleState(host)
    f0, g0, d0, vg0, gb0 = O.calc_obj_grad_clean(st, y, w_dose, w_grad)
    plan = DosePlan(dev.csr, dev.beam_ptr, dev.thr, dev.over, dev.under, store="f32", acc="f64")
    plan.set_weights(w_dose, w_grad)
    f, g = plan.eval(y)
    assert abs(f - f0) <= 1e-11 * abs(f0) and relerr(g, g0) <= 1e-11
    assert relerr(plan.dose(), d0) <= 1e-11 and relerr(plan.beamlet_grad(), gb0) <= 1e-11
    plan.close()


def test_all_zero_matrix_and_single_row_grid():
    """Degenerate inputs: D without any nonzero, and a 2 x 1 beamlet grid (the smallest the
    reference can price: with one beamlet per control point its DoseSide index runs off beamGrad)."""
    import scipy.sparse as sp
    case = make_case(V=700, nb=3, M=2, N=1, density=0.3, seed=76)
    Z = sp.csr_matrix((case.V, case.B))
    plan = DosePlan(Z, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc="f64")
    plan.set_weights(np.ones(case.B), np.ones(case.B))
    f, g = plan.eval(np.ones(case.nb))
    assert abs(f - float(np.sum(case.under * case.thr ** 2))) <= 1e-12 * max(f, 1e-300) and np.all(g == 0)
    plan.close()
    st = O.OracleState(case)
    data = vl.bind(case, store="f32", acc="f64")
    for state in (st, data):
        all_candidates(state, case)
    O.calc_obj_grad_literal(st, np.zeros(case.nb))
    vl.calcObjGrad(np.zeros(case.nb))
    got = vl.PricingProblem(0.1, 1.0, 1.0, 0.5, 2.25, 0.83, case.N, case.M, 0.5)
    want = O.pricing_problem(st, 0.1, 1.0, 1.0, 0.5, 2.25, 0.83, case.N, case.M, 0.5)
    assert got[3] == want[3] and [int(v) for v in got[1]] == want[1] and [int(v) for v in got[2]] == want[2]
    data.plan.close()


@pytest.mark.parametrize("acc,nb", [("f32", 460), ("f64", 230)])
def test_beamlet_vector_larger_than_shared_memory(acc, nb):
    """With enough beamlets x no longer fits next to the ring in shared memory: the dose kernel
    falls back to gathering x from global memory (a separate expansion kernel writes it)."""
    case = make_case(V=3000, nb=nb, M=8, N=16, density=0.002, seed=78, gastep=1)
    assert case.B * (4 if acc == "f32" else 8) > 227 * 1024
    st = O.OracleState(case)
    y, w_dose, w_grad = random_weights(case, np.random.default_rng(3), 0.6)
    f0, g0, d0, vg0, gb0 = O.calc_obj_grad_clean(st, y, w_dose, w_grad)
    plan = DosePlan(case.D, case.beam_ptr, case.thr, case.over, case.under, store="f32", acc=acc)
    assert plan.info()["launches_per_eval"] == 4
    plan.set_weights(w_dose, w_grad)
    f, g = plan.eval(y)
    tol = TOL[acc]
    assert abs(f - f0) <= tol * abs(f0) and relerr(g, g0) <= tol
    assert relerr(plan.dose(), d0) <= tol and relerr(plan.beamlet_grad(), gb0) <= tol
    plan.close()


def test_pricing_wide_rows():
    """40 beamlets per leaf row: 861 nodes per network row, dose sums in the pairwise regime."""
    case = make_case(V=5000, nb=4, M=4, N=40, density=0.02, seed=79, gastep=10)
    st = O.OracleState(case)
    data = vl.bind(case, store="f32", acc="f64")
    for state in (st, data):
        all_candidates(state, case)
    select(st, case, 1, [3, 5, 10, 2], [30, 25, 33, 18], lambda k: O.update_open_aperture(st, k))
    select(data, case, 1, [3, 5, 10, 2], [30, 25, 33, 18], vl.updateOpenAperture)
    y = np.zeros(case.nb); y[1] = 1.0
    O.calc_obj_grad_literal(st, y.copy())
    vl.calcObjGrad(y.copy())
    for C, vmax in [(0.0, 2.25), (0.05, 2.25), (0.05, 0.6)]:
        for i0 in st.notinC.loc:
            p0, l0, r0, _ = O.price_candidate(st, i0, C, 1.0, 1.0, 0.5, vmax, 0.83, case.N, case.M, 0.5)
            p, l, r, _ = vl.parallelizationPricingProblem(i0, C, 1.0, 1.0, 0.5, vmax, 0.83, case.N, case.M, 0.5)
            assert [int(v) for v in l] == l0 and [int(v) for v in r] == r0, (C, vmax, i0)
            assert abs(p - p0) <= 1e-10 * max(1.0, abs(p0))
    data.plan.close()


def test_kmedoid_costs_tree_path_and_voxel_points():
    """More than 4096 points takes the fixed-shape tree; with integer coordinates on a line the
    distances are integers, so the result is exact whatever the order."""
    from vmatwpencode_b200.plan import kmedoid_costs
    n = 5000
    pts = np.arange(n, dtype=float)
    cur = np.abs(pts - 1234.0)
    cost = kmedoid_costs(pts.reshape(-1, 1), cur)
    for c in (0, 17, 2500, 4999):
        assert cost[c] == float(np.minimum(cur, np.abs(pts - pts[c])).sum())
    rng = np.random.default_rng(8)
    P = rng.random((6000, 3))
    cur = np.sqrt(((P - P[5]) ** 2).sum(axis=1))
    cost = kmedoid_costs(P, cur)
    for c in (1, 3000):
        want = np.minimum(cur, np.sqrt(((P - P[c]) ** 2).sum(axis=1))).sum()
        assert abs(cost[c] - want) <= 1e-12 * want
    # the tiled kernel's order of additions, restated: thread t adds points t, t+256, ... left to
    # right, then a binary tree over the 256 threads; candidate count not a multiple of the tile
    n = 6003
    P = np.round(rng.random((n, 3)) * 200) / 4
    cur = np.sqrt(((P - P[11]) ** 2).sum(axis=1))
    cost = kmedoid_costs(P, cur)
    for c in (0, 7, 8, 4097, n - 1):
        s = np.zeros(n)
        for d in range(3):
            df = P[c, d] - P[:, d]
            s = s + df * df
        term = np.minimum(cur, np.sqrt(s))
        part = np.zeros(256)
        for j0 in range(0, n, 256):
            blk = term[j0:j0 + 256]
            part[:len(blk)] += blk
        o = 128
        while o:
            part[:o] += part[o:2 * o]
            o //= 2
        assert cost[c] == part[0]


def test_kmedoid_voxel_subsample_drives_the_greedy_loop():
    """BASELINE.json configs[2] in small: pick voxels by greedy k-medoid insertion over their grid
    coordinates (distance pass on the GPU), restrict the case to them, run the greedy loop on the
    subsample: same aperture sequence as the oracle on the same subsample; the sample is spatially
    even (no two sampled voxels adjacent while unsampled regions remain far larger)."""
    from vmatwpencode_b200 import kmedoids as km
    from vmatwpencode_b200.synth import voxel_coordinates, subsample_case
    case = make_case(V=6000, nb=10, M=4, N=6, density=0.03, seed=91, gastep=10)
    xyz = voxel_coordinates(case.V, case.M)
    rows = km.sample_voxels(xyz, 400, start=17)
    assert len(set(rows)) == 400 and rows[0] == 17
    # greedy insertion == brute-force argmin of the summed nearest-medoid distance, step by step
    cur = np.sqrt(((xyz - xyz[17]) ** 2).sum(axis=1))
    for nxt in rows[1:4]:
        cost = np.array([np.minimum(cur, np.sqrt(((xyz - xyz[c]) ** 2).sum(axis=1))).sum() for c in range(case.V)])
        cost[[r for r in rows[:rows.index(nxt)]]] = np.inf
        assert abs(cost[nxt] - cost.min()) <= 1e-9 * cost.min()
        cur = np.minimum(cur, np.sqrt(((xyz - xyz[nxt]) ** 2).sum(axis=1)))
    sub = subsample_case(case, rows)
    assert sub.V == 400 and sub.D.shape == (400, case.B)
    want = O.col_gen(O.OracleState(sub), C=0.01, kappasize=5, max_iter=5, clean=True)
    data = vl.bind(sub, store="f32", acc="f64")
    vl.colGen(0.01, False, 5, max_iter=5)
    assert len(want) > 0 and [(h["idx"], h["l"], h["r"]) for h in data.history] == [(h["idx"], h["l"], h["r"]) for h in want]
    data.plan.close()


@pytest.mark.parametrize("acc", ["f64", "f32"])
def test_dvh_on_device_matches_oracle(acc):
    """printresults' dose-volume histograms from the GPU-resident dose: integer bin counts, so
    exact against the oracle fed with the same dose values."""
    case = make_case(V=20_000, nb=10, M=4, N=6, density=0.02, seed=83)
    data = vl.bind(case, store="f32", acc=acc)
    all_candidates(data, case)
    for bm, (l, r) in {2: (0, 5), 7: (1, 6)}.items():
        select(data, case, bm, [l] * case.M, [r] * case.M, vl.updateOpenAperture)
    y = np.zeros(case.nb); y[2] = 3.0; y[7] = 2.0
    vl.calcObjGrad(y)
    rng = np.random.default_rng(2)
    nstructs = 6
    mask = rng.integers(1, 2 ** nstructs, size=case.V)
    mask &= ~(1 << 4)                                       # an empty structure
    mask[mask == 0] = 1
    data.fullMaskValue = mask.astype(float)
    data.numstructs = nstructs
    bins, dvh = vl.printresults(1)
    b0, m0 = O.dvh_matrix(data.currentDose, mask.astype(float), nstructs)
    assert np.array_equal(bins, b0) and np.array_equal(dvh, m0)
    assert abs(data.plan.dose_max() - data.currentDose.max()) == 0
    data.plan.close()


def test_result_bundle_round_trip(tmp_path):
    """The pickle colGen leaves behind (25 fields, :1053-1060) is readable the way picklereader.py
    reads it (positional unpacking of a list)."""
    import pickle
    case = make_case(V=4000, nb=8, M=4, N=6, density=0.03, seed=84)
    data = vl.bind(case, store="f32", acc="f64")
    vl.colGen(0.01, False, 4, max_iter=3)
    pik = str(tmp_path / "pickle-C-0.01-save.dat")
    vl.saveResults(pik, 0.01, allNames=["a", "b"])
    with open(pik, "rb") as f:
        items = pickle.load(f)
    assert len(items) == 25 and items[0] == case.nb and items[9] == case.M and items[10] == case.N
    assert np.array_equal(items[1], data.rmpres.x) and np.array_equal(items[14], data.currentDose)
    assert np.array_equal(items[19], case.thr) and items[11] == data.llist
    back = vl.loadResults(pik)
    assert back["YU"] == 10.0 / 0.83 and back["objectiveValue"] == data.objectiveValue
    data.plan.close()


@pytest.mark.parametrize("acc,tol", [("f64", 1e-11), ("f32", 1e-5)])
def test_rmp_column_cache_equals_sparse_evaluation(acc, tol):
    """solveRMC's fast path: evaluations from cached dense aperture columns equal the sparse ones."""
    case = make_case(V=30_000, nb=14, M=5, N=8, density=0.015, seed=85, ragged=True)
    st = O.OracleState(case)
    data = vl.bind(case, store="f32", acc=acc)
    for state in (st, data):
        all_candidates(state, case)
    rng = np.random.default_rng(6)
    sel = [1, 4, 5, 9, 12]
    for bm in sel:
        l = [int(rng.integers(-1, case.N - 2)) for _ in range(case.M)]
        r = [int(min(case.N, lv + rng.integers(1, 6))) for lv in l]
        select(st, case, bm, l, r, lambda k: O.update_open_aperture(st, k))
        select(data, case, bm, l, r, vl.updateOpenAperture)
    data._sync_apertures()
    data.plan.rmp_begin(sel)
    for trial in range(4):
        y = np.zeros(case.nb)
        y[sel] = rng.random(len(sel)) * 3
        f0, g0 = O.calc_obj_grad_literal(st, y.copy())
        f, g = data.plan.rmp_eval(y)
        assert abs(f - f0) <= tol * abs(f0) and relerr(g, np.asarray(g0).ravel()) <= tol
        assert relerr(data.plan.dose(), st.currentDose) <= tol
        f2, g2 = data.plan.rmp_eval(y)
        assert f2 == f and np.array_equal(g, g2)                 # reproducible
    data.plan.rmp_end()
    f3, g3 = data.plan.eval(y)                                    # the sparse path still works afterwards
    assert abs(f3 - f0) <= tol * abs(f0)
    data.plan.close()


def test_greedy_loop_same_with_and_without_column_cache():
    z, case = load_golden("colgen_B")
    seqs = []
    for flag in (True, False):
        vl.use_column_cache = flag
        try:
            data = vl.bind(case, store="f32", acc="f64")
            vl.colGen(float(z["C"]), False, int(z["kappasize"]), max_iter=int(z["max_iter"]))
            seqs.append([(h["idx"], tuple(h["l"]), tuple(h["r"])) for h in data.history])
            data.plan.close()
        finally:
            vl.use_column_cache = False
    assert seqs[0] == seqs[1] == [(int(i), tuple(l), tuple(r)) for i, l, r in zip(z["idx"], z["l"], z["r"])]


# ---- the two earlier generations of the driver functions (vmatwpencode_b200/variants.py) ------------------------
def _variant_api(name):
    from vmatwpencode_b200 import variants
    return variants.short if "short" in name else variants.classic


@pytest.mark.parametrize("name", ["variant_short_A", "variant_short_B", "variant_classic_A", "variant_classic_B",
                                  "variant_classic_C"])
def test_variant_greedy_runs_against_reference_fixture(name):
    """greedyVMATshort.py / greedyVMAT.py semantics on the GPU path against runs recorded from the reference's own
    functions: control points, leaf vectors (bit-exact, including the trailing sink entry and the odd paths of
    the shifted relaxation window), prices, intensities, objective."""
    z, case = load_golden(name)
    g = _variant_api(name)
    g.gastep = int(z["gastep"])
    data = vl.bind(case, store="f32", acc="f64")
    g.colGen(float(z["C"]), max_iter=int(z["max_iter"]))
    off = np.concatenate([[0], np.cumsum(z["l_len"])])
    assert [h["idx"] for h in data.history] == list(z["idx"])
    for k, h in enumerate(data.history):
        assert h["l"] == list(z["l_flat"][off[k]:off[k + 1]]) and h["r"] == list(z["r_flat"][off[k]:off[k + 1]])
        assert abs(h["pstar"] - z["pstar"][k]) <= 1e-9 * abs(z["pstar"][k])
        assert abs(h["fun"] - z["fun"][k]) <= 1e-9 * abs(z["fun"][k])
        assert np.abs(h["x"] - z["x"][k]).max() <= 1e-6 * max(1.0, np.abs(z["x"][k]).max())
    # every remaining candidate priced in the final state, one by one (PPsubroutine) and batched (PricingProblem)
    data.calcDose()
    data.calcGradientandObjValue()
    offf = np.concatenate([[0], np.cumsum(z["final_len"])])
    C = float(z["C"])
    for k, i in enumerate(z["final_idx"]):
        predec, succ = g._neighbours(int(i))
        if g.name == "short":
            am = ap = 60
        else:
            ap = (succ - int(i)) * g.gastep if type(succ) is not list else np.inf
            am = (int(i) - predec) * g.gastep if type(predec) is not list else np.inf
        p, l, r = g.PPsubroutine(C, 1.0, 1.0, 0.5, am, ap, g.vmax, g.speedlim, predec, succ, case.N, case.M, int(i))
        assert abs(p - z["final_p"][k]) <= 1e-9 * max(abs(z["final_p"][k]), 1e-300)
        assert [int(v) for v in l] == list(z["final_l_flat"][offf[k]:offf[k + 1]])
        assert [int(v) for v in r] == list(z["final_r_flat"][offf[k]:offf[k + 1]])
    data.plan.close()


@pytest.mark.parametrize("gen", ["short", "classic"])
def test_variant_pricing_random_states_vs_oracle(gen):
    """Per-candidate pricing of both earlier generations in random selected states (random leaf vectors of the
    neighbours, loose and tight leaf speeds) against oracle/vmat_oracle_variants.py; an exception of the oracle
    (the reference's NameError / UnboundLocalError / ValueError paths) must be raised by the product as well."""
    from oracle import vmat_oracle_variants as V
    from vmatwpencode_b200 import variants
    g = getattr(variants, gen)
    var = V.VARIANTS[gen]
    rng = np.random.default_rng(91)
    for trial in range(6):
        case = make_case(V=3000, nb=10, M=5 + trial % 3, N=6 + trial % 4, density=0.03, seed=300 + trial,
                         gastep=[10, 2, 1][trial % 3], ragged=bool(trial & 1))
        gastep = case.angles[1] - case.angles[0]
        g.gastep = gastep
        st = V.VariantState(case)
        data = vl.bind(case, store="f32", acc="f64")
        sel = sorted(rng.choice(case.nb, size=4, replace=False).tolist())
        st.caligraphicC, st.notinC = list(sel), [i for i in range(case.nb) if i not in sel]
        data.caligraphicC, data.notinC = list(sel), list(st.notinC)
        for i in sel:
            ll, rr = [], []
            for m in range(case.M):              # leaves inside the row's beamlets (outside, the reference's index arithmetic
                vb = O.valid_beamlets(st, m, i)  # leaves the control point's vector and raises IndexError)
                l = int(rng.integers(vb.min() - 1, vb.max()))
                ll.append(l)
                rr.append(int(min(l + rng.integers(1, 4), vb.max() + 1)))
            st.llist[i] = data.llist[i] = ll
            st.rlist[i] = data.rlist[i] = rr
            st.openApertureMaps[i], st.diagmakers[i] = V.update_open_aperture(st, i)
            data.openApertureMaps[i], data.diagmakers[i] = g.updateOpenAperture(i)
            assert np.array_equal(st.openApertureMaps[i], data.openApertureMaps[i])
            assert np.array_equal(st.diagmakers[i], data.diagmakers[i])
            data.strengths[i] = [1.0] * len(data.openApertureMaps[i])
        y = np.zeros(case.nb)
        y[sel] = rng.random(4) * 2
        f0, g0 = V.calc_obj_grad(st, y.copy())
        f1, g1 = g.calcObjGrad(y.copy())
        assert abs(f1 - f0) <= 1e-11 * abs(f0) and relerr(g1, g0) <= 1e-11
        for vmax in (var.vmax, 0.05):
            for i in st.notinC:
                predec, succ = V.neighbours(st, var, i, gastep)
                if gen == "short":
                    am = ap = 60
                else:
                    ap = (succ - i) * gastep if not isinstance(succ, list) else np.inf
                    am = (i - predec) * gastep if not isinstance(predec, list) else np.inf
                args = (0.3, 1.0, 1.0, 0.5, am, ap, vmax, var.speedlim, predec, succ, case.N, case.M, i)
                try:
                    want = V.pp_subroutine(st, var, *args)
                except (NameError, UnboundLocalError, ValueError) as e:
                    with pytest.raises(type(e)):
                        g.PPsubroutine(*args)
                    continue
                p, l, r = g.PPsubroutine(*args)
                assert [int(v) for v in l] == want[1] and [int(v) for v in r] == want[2], (trial, i, vmax)
                assert abs(p - want[0]) <= 1e-9 * max(abs(want[0]), 1e-300)
        data.plan.close()


def test_greedy_loop_config1_short_semantics_vs_oracle():
    """BASELINE.json configs[0] as it names it: greedyVMATshort.py on the 20k-voxel x 2016-beamlet x 36-control-
    point phantom - the first-generation greedy loop to completion, same aperture sequence as the oracle."""
    from oracle import vmat_oracle_variants as V
    from vmatwpencode_b200 import variants
    case = make_config("cfg1")
    want = V.col_gen(V.VariantState(case), V.SHORT, C=1, max_iter=12)
    data = vl.bind(case, store="f32", acc="f64")
    variants.short.colGen(1, max_iter=12)
    assert len(want) > 0 and [h["idx"] for h in data.history] == [w["idx"] for w in want]
    for h, w in zip(data.history, want):
        assert h["l"] == w["l"] and h["r"] == w["r"]
        assert abs(h["fun"] - w["fun"]) <= 1e-8 * abs(w["fun"])
    data.plan.close()


def test_lean_rmp_loop_equals_scipy_minimize():
    """solveRMC through the lean loop over scipy's compiled L-BFGS-B routine and through scipy.optimize.minimize:
    same iterates, bit for bit, and the same greedy sequences."""
    case = make_case(V=20_000, nb=24, M=6, N=8, density=0.01, seed=88, gastep=10)
    runs = {}
    for fast in (True, False):
        vl.fast_rmp = fast
        try:
            data = vl.bind(case, store="f32", acc="f64")
            vl.colGen(0.01, False, 8, max_iter=8)
            runs[fast] = [(h["idx"], h["l"], h["r"], h["fun"], h["nit"], h["nfev"], h["x"].tobytes()) for h in data.history]
            last = (data.objectiveValue, np.asarray(data.aperturegradient).ravel().copy(), np.array(data.currentIntensities))
            runs[(fast, "state")] = last
            data.plan.close()
        finally:
            vl.fast_rmp = True
    assert len(runs[True]) > 0 and runs[True] == runs[False]
    a, b = runs[(True, "state")], runs[(False, "state")]
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_greedy_loop_whole_circle_branch_with_the_reference_kappa():
    """colGen(C, WholeCircle=True, ...) (greedyVMATsampling.py:959-966): the candidates come from the module-level
    `kappa`, for which the reference hard-codes its 178-entry k-medoid ordering of the 2-degree control points (:30,
    recorded in tests/golden/kmedoids.npz).  178 control points, gantry step 2: neighbouring apertures bound the
    leaf ranges (travel 10.8 beamlets)."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmedoids.npz"))
    kappa = [int(v) for v in z["kappa178"]]
    case = make_case(V=8000, nb=178, M=4, N=14, density=0.004, seed=178, gastep=2)
    want = O.col_gen(O.OracleState(case), C=0.02, kappasize=16, max_iter=7, clean=True, kappa0=kappa)
    data = vl.bind(case, store="f32", acc="f64")
    vl.kappa = list(kappa)
    try:
        vl.colGen(0.02, True, 16, max_iter=7)
    finally:
        vl.kappa = []
    assert len(want) > 0 and [h["idx"] for h in data.history] == [w["idx"] for w in want]
    for h, w in zip(data.history, want):
        assert h["l"] == w["l"] and h["r"] == w["r"]
    data.plan.close()


def test_kmedoid_resident_session_with_candidate_subset():
    """Voxel sampling with the candidates restricted to a subset and the points resident on the device between the
    insertion steps (vmat_kmedoid_open / insert / costs_resident): step by step the same medoids as a brute-force
    numpy argmin over the candidate set, for a small (sequential-sum kernel) and a large (tiled kernel) point set."""
    from vmatwpencode_b200 import kmedoids as km
    rng = np.random.default_rng(23)
    for n, stride, k in ((900, 7, 12), (9000, 31, 10)):
        P = np.round(rng.random((n, 3)) * 40)
        cand = np.arange(0, n, stride)
        start = int(cand[len(cand) // 2])
        got = km.sample_voxels(P, k, start=start, candidates=cand)
        kap = [start]
        cur = np.sqrt(((P - P[start]) ** 2).sum(axis=1))
        for _ in range(k - 1):
            cost = np.array([np.minimum(cur, np.sqrt(((P - P[c]) ** 2).sum(axis=1))).sum() if c not in kap else np.inf
                             for c in cand])
            best = int(cand[int(np.argmin(cost))])
            assert abs(cost.min() - np.minimum(cur, np.sqrt(((P - P[got[len(kap)]]) ** 2).sum(axis=1))).sum()) <= 1e-9 * cost.min()
            kap.append(best)
            cur = np.minimum(cur, np.sqrt(((P - P[best]) ** 2).sum(axis=1)))
        assert got == kap
