"""Live pin: oracle/vmat_oracle.py against the reference's own functions, AST-extracted
from /root/reference at run time (oracle/ref_extract.py).  Skipped where the reference is not
mounted (the GPU box); tests/test_golden_oracle.py carries the same pin as fixtures."""
import warnings

import numpy as np
import pytest

from oracle import ref_extract as R
from oracle import vmat_oracle as O
from vmatwpencode_b200.synth import make_case

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="reference not mounted")


@pytest.mark.parametrize("cfg", [
    dict(V=1200, nb=8, M=4, N=5, density=0.05, seed=41, gastep=10, C=0.02, kappa=4),
    dict(V=1500, nb=10, M=3, N=8, density=0.05, seed=42, gastep=2, C=0.0, kappa=5, ragged=True),
    dict(V=1000, nb=20, M=3, N=5, density=0.06, seed=43, gastep=1, C=0.3, kappa=6),
])
def test_colgen_bitwise(cfg):
    cfg = dict(cfg)
    C, kappasize = cfg.pop("C"), cfg.pop("kappa")
    case = make_case(**cfg)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = R.ref_colgen(R.reference_namespace(case), case, C=C, kappasize=kappasize, max_iter=8)
    got = O.col_gen(O.OracleState(case), C=C, kappasize=kappasize, max_iter=8)
    assert len(ref) == len(got) and len(ref) > 0
    for a, b in zip(ref, got):
        assert a["idx"] == b["idx"] and a["l"] == b["l"] and a["r"] == b["r"]
        assert a["pstar"] == b["pstar"] and a["fun"] == b["fun"]
        assert np.array_equal(a["x"], b["x"])


def test_update_open_aperture_fractional_leaves():
    case = make_case(V=600, nb=4, M=4, N=9, density=0.08, seed=44, gastep=10, ragged=True)
    ns = R.reference_namespace(case)
    st = O.OracleState(case)
    rng = np.random.default_rng(3)
    for trial in range(40):
        beam = int(rng.integers(0, case.nb))
        l = [float(rng.integers(-1, case.N - 1)) + (rng.choice([0.0, 0.25, 0.5, 0.995]) if trial % 2 else 0.0)
             for _ in range(case.M)]
        r = [min(case.N + 1.0, lv + float(rng.integers(0, 5)) + (rng.choice([0.0, 0.3, 0.75, 0.005]) if trial % 3 else 0.0))
             for lv in l]
        for state in (ns["data"], st):
            state.llist[beam] = list(l)
            state.rlist[beam] = list(r)
        with R.quiet():
            a = ns["updateOpenAperture"](beam)
        b = O.update_open_aperture(st, beam)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and list(a[2]) == list(b[2])


def test_kmedoids_functions():
    ns = R.extract("kmedoids")
    with R.quiet():
        a = ns["givemewholelist"]([2, 9], range(0, 17))
        d = ns["distancemin"]([2, 9, 4], range(0, 17))
    assert O.km_givemewholelist([2, 9], list(range(17))) == a
    assert O.km_distancemin([2, 9, 4], list(range(17))) == d


@pytest.mark.parametrize("module,var", [("greedyVMATshort", "short"), ("greedyVMAT", "classic")])
def test_variant_oracle_against_live_reference(module, var):
    """oracle/vmat_oracle_variants.py against the functions of greedyVMATshort.py / greedyVMAT.py executed here:
    whole greedy runs, bit for bit (also with a tight gantry step for the classic generation, where neighbouring
    apertures bound the leaf ranges)."""
    import warnings
    from oracle import vmat_oracle_variants as V
    for kw, C in ((dict(V=1200, nb=5, M=4, N=6, density=0.05, seed=31, gastep=60), 0.5),
                  (dict(V=2000, nb=7, M=6, N=7, density=0.04, seed=32, gastep=2, ragged=True), 0.05)):
        case = make_case(**kw)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = R.variant_namespace(case, module, gastep=kw["gastep"])
            try:
                ref = R.variant_colgen(ns, case, module, C=C, max_iter=5)
            except UnboundLocalError as e:       # no path into the sink at or below its initial 0.0: p is never assigned there
                ref = type(e)
        try:
            got = V.col_gen(V.VariantState(case), V.VARIANTS[var], C=C, gastep=kw["gastep"], max_iter=5)
        except UnboundLocalError as e:
            got = type(e)
        if isinstance(ref, type) or isinstance(got, type):
            assert ref is got
            continue
        assert len(ref) == len(got) and len(ref) > 0
        for a, b in zip(ref, got):
            assert a["idx"] == b["idx"] and a["l"] == b["l"] and a["r"] == b["r"]
            assert a["pstar"] == b["pstar"] and a["fun"] == b["fun"] and np.array_equal(a["x"], b["x"])
