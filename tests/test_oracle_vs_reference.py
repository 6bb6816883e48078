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
