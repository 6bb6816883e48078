"""The oracle (oracle/vmat_oracle.py) against the fixtures frozen from the reference's own
functions (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import vmat_oracle as O
from helpers import load_golden, oracle_state_from_golden, relerr


@pytest.mark.parametrize("name", ["evalprice_A", "evalprice_B", "evalprice_C", "evalprice_D"])
def test_eval_matches_reference_bitwise(name):
    z, case = load_golden(name)
    st = oracle_state_from_golden(z, case)
    for beam in z["selected"]:
        beam = int(beam)
        assert np.array_equal(st.openApertureMaps[beam], z[f"sel{beam}_oam"])
        assert np.array_equal(st.diagmakers[beam], z[f"sel{beam}_diag"])
        assert np.array_equal(np.asarray(st.strengths[beam], dtype=float), z[f"sel{beam}_strength"])
    f, g = O.calc_obj_grad_literal(st, np.array(z["y"]))
    assert f == float(z["f"])
    assert np.array_equal(np.asarray(g).ravel(), z["g"])
    assert np.array_equal(st.currentDose, z["dose"])
    assert np.array_equal(st.voxelgradient, z["vg"])
    # the global ("clean") form the kernels implement agrees to rounding
    f2, g2, d2, vg2, gb2 = O.calc_obj_grad_clean(st, np.array(z["y"]))
    assert abs(f2 - float(z["f"])) <= 1e-13 * abs(float(z["f"]))
    assert relerr(g2, z["g"]) <= 1e-13
    assert relerr(d2, z["dose"]) <= 1e-13


@pytest.mark.parametrize("name", ["evalprice_A", "evalprice_B", "evalprice_C", "evalprice_D"])
def test_pricing_matches_reference_bitwise(name):
    z, case = load_golden(name)
    st = oracle_state_from_golden(z, case)
    O.calc_obj_grad_literal(st, np.array(z["y"]))
    for k in range(int(z["n_price"])):
        C, C2, C3, b, vmax, speedlim, bw = [float(v) for v in z[f"price{k}_params"]]
        for n, idx in enumerate(z[f"price{k}_idx"]):
            p, l, r, i = O.price_candidate(st, int(idx), C, C2, C3, b, vmax, speedlim, case.N, case.M, bw)
            assert p == z[f"price{k}_p"][n], (name, k, idx)
            assert l == list(z[f"price{k}_l"][n]) and r == list(z[f"price{k}_r"][n]), (name, k, idx)


def test_fallback_midpoints_are_exercised():
    """The tight leaf-speed settings of the fixtures must reach the fractional-midpoint
    branch (quirk Q7), otherwise that code is unpinned."""
    z, case = load_golden("evalprice_B")
    st = oracle_state_from_golden(z, case)
    hit = 0
    for k in range(int(z["n_price"])):
        C, C2, C3, b, vmax, speedlim, bw = [float(v) for v in z[f"price{k}_params"]]
        for idx in z[f"price{k}_idx"]:
            predec, succ, am, ap = O.neighbours(st, int(idx))
            if isinstance(predec, list) or isinstance(succ, list):
                continue
            tm, tp = vmax * (am / speedlim) / bw, vmax * (ap / speedlim) / bw
            win = O.leaf_window(case.N, case.M, st.llist[predec], st.llist[succ], st.rlist[predec],
                                st.rlist[succ], tm, tp, am, ap)
            for row in win:
                for l, rs in row:
                    if not float(l).is_integer() or any(not float(r).is_integer() for r in rs):
                        hit += 1
    assert hit > 0


@pytest.mark.parametrize("name", ["colgen_A", "colgen_B", "colgen_C", "colgen_E"])
def test_colgen_sequence_matches_reference(name):
    """Greedy sequences recorded from the reference's own PricingProblem / updateOpenAperture /
    solveRMC (colgen_E: Laptop variant with the elimination phase switched on)."""
    z, case = load_golden(name)
    st = O.OracleState(case)
    elim = bool(z["elimination"])
    hist = O.col_gen(st, C=float(z["C"]), kappasize=int(z["kappasize"]), max_iter=int(z["max_iter"]), elimination=elim)
    assert [h["idx"] for h in hist] == list(z["idx"])
    assert [len(h["removed"]) for h in hist] == list(z["removed_count"])
    assert [k for h in hist for k in h["removed"]] == list(z["removed_flat"])
    if elim:
        assert int(z["removed_count"].sum()) > 0, "fixture no longer exercises the elimination phase"
    if "kelly" in z.files:      # Kelly measure of the control points still selected at the end
        last = {}
        for h in hist:
            last[h["idx"]] = h["kelly"]
            for k in h["removed"]:
                last.pop(k, None)
        assert sorted(last) == list(z["kelly_idx"])
        assert [last[k] for k in sorted(last)] == list(z["kelly"])
    for h, l, r, p, x, fun in zip(hist, z["l"], z["r"], z["pstar"], z["x"], z["fun"]):
        assert h["l"] == list(l) and h["r"] == list(r)
        assert h["pstar"] == p and h["fun"] == fun
        assert np.array_equal(h["x"], x)
    # the global-form evaluation drives the same greedy sequence
    st2 = O.OracleState(case)
    hist2 = O.col_gen(st2, C=float(z["C"]), kappasize=int(z["kappasize"]), max_iter=int(z["max_iter"]), clean=True,
                      elimination=elim)
    assert [h["idx"] for h in hist2] == list(z["idx"])
    for h, l, r in zip(hist2, z["l"], z["r"]):
        assert h["l"] == list(l) and h["r"] == list(r)


def test_kmedoids_known_answer_and_reference_runs():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "kmedoids.npz"))
    for n in (23, 40, 31):
        start = [int(v) for v in z[f"start_n{n}"]]
        assert O.km_givemewholelist(start, list(range(n))) == list(z[f"order_n{n}"])
        assert O.km_distancemin(start, list(range(n))) == float(z[f"dmin_n{n}"])
    # first steps of the 178-angle ordering embedded in the reference (greedyVMAT.py:28)
    kappa = list(range(6, 178, 11))
    for _ in range(4):
        kappa = O.km_insertnewelement(kappa, list(range(178)))
    assert kappa == list(z["kappa178"][:len(kappa)])


def test_numpy_sum_order_restatement():
    rng = np.random.default_rng(5)
    for n in range(0, 65):
        for _ in range(20):
            a = rng.standard_normal(n) * 10.0 ** rng.integers(-4, 4, size=n)
            if n:
                assert O.numpy_pairwise_sum(a) == a[np.arange(n)].sum()


def test_global_form_equals_literal_when_open_map_has_duplicates():
    """Ragged beamlet rows with a left leaf left of the row's first beamlet make the
    reference's open map restart at index 0 and list beamlets twice; the global form (what the
    kernels implement) must deposit them twice as well."""
    from vmatwpencode_b200.synth import make_case
    case = make_case(V=2000, nb=6, M=6, N=12, density=0.05, seed=80, gastep=2, ragged=True)
    st = O.OracleState(case)
    rng = np.random.default_rng(0)
    dup = 0
    y = np.zeros(case.nb)
    for bm in (1, 3, 4):
        st.caligraphicC.insertAngle(bm, case.angles[bm])
        st.llist[bm] = [int(rng.integers(-1, case.N - 2)) for _ in range(case.M)]
        st.rlist[bm] = [int(min(case.N, lv + rng.integers(1, 6))) for lv in st.llist[bm]]
        st.openApertureMaps[bm], st.diagmakers[bm], st.strengths[bm] = O.update_open_aperture(st, bm)
        dup += len(st.openApertureMaps[bm]) - len(set(st.openApertureMaps[bm].tolist()))
        y[bm] = 0.5 + rng.random()
    assert dup > 0, "case no longer exercises duplicate open-map entries"
    f1, g1 = O.calc_obj_grad_literal(st, y.copy())
    d1 = st.currentDose.copy()
    f2, g2, d2, vg2, gb2 = O.calc_obj_grad_clean(st, y.copy())
    assert relerr(d2, d1) <= 1e-13 and abs(f2 - f1) <= 1e-13 * abs(f1) and relerr(g2, g1) <= 1e-13


def test_dvh_matches_reference_printresults():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "dvh.npz"))
    b, m = O.dvh_matrix(z["dose"], z["full_mask"].astype(float), int(z["numstructs"]))
    assert np.array_equal(b, z["bin_center"]) and np.array_equal(m, z["dvh_matrix"])
    assert np.all(m[5] == 0)          # the empty structure keeps a zero row


# ---- the two earlier generations (greedyVMATshort.py, greedyVMAT.py): oracle/vmat_oracle_variants.py ----------
def _variant_history(z):
    off = np.concatenate([[0], np.cumsum(z["l_len"])])
    return [(int(z["idx"][k]), list(z["l_flat"][off[k]:off[k + 1]]), list(z["r_flat"][off[k]:off[k + 1]]),
             float(z["pstar"][k]), z["x"][k], float(z["fun"][k])) for k in range(len(z["idx"]))]


@pytest.mark.parametrize("name", ["variant_short_A", "variant_short_B", "variant_classic_A", "variant_classic_B",
                                  "variant_classic_C"])
def test_variant_oracle_reproduces_reference_runs_bitwise(name):
    """Greedy runs of greedyVMATshort.py / greedyVMAT.py recorded from the reference's own functions
    (oracle/make_golden.py: variant_fixture): same control points, leaf vectors (including the trailing sink
    entry and the paths the shifted window produces), prices, intensities and objective values, bit for bit."""
    from oracle import vmat_oracle_variants as V
    z, case = load_golden(name)
    var = V.SHORT if "short" in name else V.CLASSIC
    st = V.VariantState(case)
    got = V.col_gen(st, var, C=float(z["C"]), gastep=int(z["gastep"]), max_iter=int(z["max_iter"]))
    want = _variant_history(z)
    assert len(got) == len(want)
    for h, (idx, l, r, p, x, fun) in zip(got, want):
        assert h["idx"] == idx and h["l"] == l and h["r"] == r
        assert h["pstar"] == p and h["fun"] == fun and np.array_equal(h["x"], x)
    # per-candidate prices of the final state
    V.calc_obj_grad(st, st.currentIntensities)
    off = np.concatenate([[0], np.cumsum(z["final_len"])])
    gastep = int(z["gastep"])
    for k, i in enumerate(z["final_idx"]):
        predec, succ = V.neighbours(st, var, int(i), gastep)
        if var.rows == "short":
            am = ap = 60
        else:
            ap = (succ - int(i)) * gastep if not isinstance(succ, list) else np.inf
            am = (int(i) - predec) * gastep if not isinstance(predec, list) else np.inf
        p, l, r = V.pp_subroutine(st, var, float(z["C"]), 1.0, 1.0, 0.5, am, ap, var.vmax, var.speedlim, predec, succ,
                                  case.N, case.M, int(i))
        assert p == z["final_p"][k]
        assert l == list(z["final_l_flat"][off[k]:off[k + 1]]) and r == list(z["final_r_flat"][off[k]:off[k + 1]])
