"""Shared test helpers: golden fixtures, putting an oracle state / the product's ``data``
into a given aperture configuration."""
import os

import numpy as np

from oracle import vmat_oracle as O
from vmatwpencode_b200.synth import case_from_arrays

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, case_from_arrays(z)


def leaf_list(a):
    """Leaf vectors as the reference holds them: python ints where integral, floats otherwise."""
    return [int(v) if float(v).is_integer() else float(v) for v in a]


def all_candidates(state, case):
    state.notinC.loc.clear(); state.notinC.angle.clear()
    state.caligraphicC.loc.clear(); state.caligraphicC.angle.clear()
    for i in range(case.nb):
        state.notinC.insertAngle(i, case.angles[i])


def select(state, case, beam, l, r, update):
    """Move control point `beam` into the selected set with leaf vectors l, r."""
    if beam in state.notinC.loc:
        state.notinC.removeIndex(beam)
    state.caligraphicC.insertAngle(beam, case.angles[beam])
    state.llist[beam] = leaf_list(l)
    state.rlist[beam] = leaf_list(r)
    out = update(beam)
    state.openApertureMaps[beam], state.diagmakers[beam], state.strengths[beam] = out
    return out


def oracle_state_from_golden(z, case):
    st = O.OracleState(case)
    all_candidates(st, case)
    for beam in z["selected"]:
        beam = int(beam)
        select(st, case, beam, z[f"sel{beam}_l"], z[f"sel{beam}_r"], lambda k: O.update_open_aperture(st, k))
    return st


def relerr(a, b):
    a = np.asarray(a, dtype=float).ravel(); b = np.asarray(b, dtype=float).ravel()
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)
