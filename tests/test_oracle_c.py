"""oracle/vmat_oracle.c (the C restatement used as checker and CPU baseline) against the
numpy oracle.  CPU only."""
import os

import numpy as np
import pytest

from oracle import vmat_oracle as O
from vmatwpencode_b200.synth import make_case

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "oracle", "liboracle.so")


@pytest.mark.skipif(not os.path.isfile(LIB), reason="oracle/liboracle.so not built (run __graft_entry__.build())")
@pytest.mark.parametrize("threads", [1, 3])
def test_c_oracle_equals_numpy_oracle(threads):
    from oracle.c_oracle import COracle
    case = make_case(V=4000, nb=9, M=4, N=7, density=0.03, seed=51, gastep=10, ragged=True)
    st = O.OracleState(case)
    rng = np.random.default_rng(1)
    w_dose = (rng.random(case.B) > 0.5) * rng.random(case.B)
    w_grad = (rng.random(case.B) > 0.5) * 1.0
    y = rng.random(case.nb) * 4
    f, g, d, vg, gb = O.calc_obj_grad_clean(st, y, w_dose, w_grad)
    co = COracle(case)
    f2, g2 = co.eval(y, w_dose, w_grad, threads=threads)
    assert f2 == f
    assert np.array_equal(co.d, d) and np.array_equal(co.vg, vg)
    assert np.abs(co.gb - gb).max() <= 1e-14 * np.abs(gb).max()
    assert np.abs(g2 - g).max() <= 1e-14 * np.abs(g).max()
