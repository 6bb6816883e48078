"""TEST INFRASTRUCTURE - ctypes wrapper of oracle/liboracle.so (oracle/vmat_oracle.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _lib.oracle_eval.restype = C.c_double
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


class COracle:
    """float64 CPU evaluation of one case (both orientations of D held like scipy would)."""

    def __init__(self, case):
        self.lib = load()
        Dr = case.D.tocsr(); Dr.sort_indices()
        Dc = case.D.tocsc(); Dc.sort_indices()
        self.V, self.B, self.nb = case.V, case.B, case.nb
        self.beam_ptr = np.ascontiguousarray(case.beam_ptr, dtype=np.int32)
        self.a = [np.ascontiguousarray(Dr.indptr, dtype=np.int64), np.ascontiguousarray(Dr.indices, dtype=np.int32),
                  np.ascontiguousarray(Dr.data, dtype=np.float64),
                  np.ascontiguousarray(Dc.indptr, dtype=np.int64), np.ascontiguousarray(Dc.indices, dtype=np.int32),
                  np.ascontiguousarray(Dc.data, dtype=np.float64),
                  np.ascontiguousarray(case.thr), np.ascontiguousarray(case.over), np.ascontiguousarray(case.under)]
        self.x = np.zeros(self.B); self.d = np.zeros(self.V); self.vg = np.zeros(self.V)
        self.ft = np.zeros(self.V); self.gb = np.zeros(self.B); self.g = np.zeros(self.nb)

    @classmethod
    def from_arrays(cls, V, B, nb, beam_ptr, vrowptr, vcol, vval, browptr, bcol, bval, thr, over, under):
        """The same checker on CSR arrays that already exist in both orientations (e.g. copied back from a
        device-generated case): no scipy conversion, values widened to float64."""
        self = cls.__new__(cls)
        self.lib = load()
        self.V, self.B, self.nb = int(V), int(B), int(nb)
        self.beam_ptr = np.ascontiguousarray(beam_ptr, dtype=np.int32)
        self.a = [np.ascontiguousarray(vrowptr, dtype=np.int64), np.ascontiguousarray(vcol, dtype=np.int32),
                  np.ascontiguousarray(vval, dtype=np.float64),
                  np.ascontiguousarray(browptr, dtype=np.int64), np.ascontiguousarray(bcol, dtype=np.int32),
                  np.ascontiguousarray(bval, dtype=np.float64),
                  np.ascontiguousarray(thr, dtype=np.float64), np.ascontiguousarray(over, dtype=np.float64),
                  np.ascontiguousarray(under, dtype=np.float64)]
        self.x = np.zeros(self.B); self.d = np.zeros(self.V); self.vg = np.zeros(self.V)
        self.ft = np.zeros(self.V); self.gb = np.zeros(self.B); self.g = np.zeros(self.nb)
        return self

    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def eval(self, y, w_dose, w_grad, threads=0):
        y = np.ascontiguousarray(y, dtype=np.float64)
        w_dose = np.ascontiguousarray(w_dose, dtype=np.float64)
        w_grad = np.ascontiguousarray(w_grad, dtype=np.float64)
        p = lambda a: C.c_void_p(a.ctypes.data)
        f = self.lib.oracle_eval(C.c_int64(self.V), C.c_int64(self.B), C.c_int32(self.nb), p(self.beam_ptr),
                                 *[p(a) for a in self.a], p(y), p(w_dose), p(w_grad),
                                 p(self.x), p(self.d), p(self.vg), p(self.ft), p(self.gb), p(self.g),
                                 C.c_int(threads))
        return f, self.g.copy()
