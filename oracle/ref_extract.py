"""TEST INFRASTRUCTURE - run the reference's OWN hot-path functions on a synthetic case.

Only usable where /root/reference is mounted (the build container).  Nothing
from the reference is copied into this repository: the wanted top-level
``def``/``class`` nodes are parsed out of the reference files with ``ast`` at
run time and executed in a scratch namespace that is pre-seeded with synthetic
globals.  Used by oracle/make_golden.py to produce the
fixtures under tests/golden/ and by tests/test_oracle_vs_reference.py (skipped
when the reference is absent) to pin oracle/vmat_oracle.py.

Product code must never import this module.
"""
from __future__ import annotations

import ast
import contextlib
import io
import math
import os
import random
import sys
import time

import numpy as np
from scipy import sparse
from scipy.optimize import minimize

REFERENCE_ROOT = os.environ.get("VMAT_REFERENCE_ROOT", "/root/reference")

_HOT = ["region", "apertureList", "vmat_class", "calcObjGrad", "fvalidbeamlets", "PPsubroutine",
        "parallelizationPricingProblem", "PricingProblem", "solveRMC", "updateOpenAperture"]
WANTED = {
    "greedyVMATsampling": _HOT,
    "greedyVMATsamplingLaptop": _HOT,
    # the two earlier generations: same function names, plain-list candidate sets
    "greedyVMAT": ["region", "vmat_class", "calcObjGrad", "fvalidbeamlets", "PPsubroutine", "parallelizationPricingProblem",
                   "PricingProblem", "solveRMC", "updateOpenAperture"],
    "greedyVMATshort": ["region", "vmat_class", "calcObjGrad", "PPsubroutine", "PricingProblem", "solveRMC",
                        "updateOpenAperture"],
    "kmedoids": ["distancemin", "listdiff", "insertnewelement", "givemewholelist"],
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "greedyVMATsampling.py"))


def _numpy_shim():
    """numpy with the three aliases the reference still uses (np.int, np.float, np.NAN)."""
    import types
    shim = types.ModuleType("numpy_shim")
    shim.__dict__.update(np.__dict__)
    shim.int = int
    shim.float = float
    shim.NAN = np.nan
    return shim


def extract(module: str, names=None) -> dict:
    """exec the wanted top-level definitions of /root/reference/<module>.py; return the namespace."""
    path = os.path.join(REFERENCE_ROOT, module + ".py")
    tree = ast.parse(open(path).read())
    names = set(names or WANTED[module])
    keep = [n for n in tree.body
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    if module == "kmedoids":
        keep = [n for n in tree.body if isinstance(n, ast.ImportFrom) and n.module == "scipy.spatial"] + keep
    from functools import partial

    class SerialPool:
        """Stand-in for multiprocessing.Pool so that the reference's own PricingProblem runs here
        (its pool is created under ``if __name__ == '__main__'``, hence __name__ below)."""

        def __init__(self, processes=None):
            pass

        def map(self, fn, it):
            return [fn(x) for x in it]

        def close(self):
            pass

        def join(self):
            pass

    ns = {"np": _numpy_shim(), "sparse": sparse, "math": math, "time": time, "random": random,
          "sys": sys, "minimize": minimize, "__name__": "__main__", "Pool": SerialPool, "partial": partial,
          "numcores": 1}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


@contextlib.contextmanager
def quiet():
    """The reference prints a lot; swallow it."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def reference_namespace(case, module: str = "greedyVMATsampling") -> dict:
    """Namespace with the reference's functions bound to `case` through the globals
    they read: ``data``, ``DlistT``, ``quadHelperThresh/Over/Under`` (Appendix A step 3)."""
    ns = extract(module)
    data = ns["vmat_class"]()
    data.numvoxels = case.V
    data.numbeams = case.nb
    data.xinter = np.array(case.xinter)
    data.yinter = np.array(case.yinter)
    data.xdirection = [np.array(a) for a in case.xdirection]
    data.ydirection = [np.array(a) for a in case.ydirection]
    data.Dlist = case.Dlist()
    data.llist = [[-1] * case.M for _ in range(case.nb)]
    data.rlist = [[case.N + 1] * case.M for _ in range(case.nb)]
    data.openApertureMaps = [[] for _ in range(case.nb)]
    data.diagmakers = [[] for _ in range(case.nb)]
    data.strengths = [[] for _ in range(case.nb)]
    data.currentIntensities = np.zeros(case.nb)
    data.pointtoAngle = list(case.angles)
    data.notinC = ns["apertureList"]()
    data.caligraphicC = ns["apertureList"]()
    data.listIndexofAperturesRemovedEachStep = []
    ns["data"] = data
    ns["DlistT"] = [d.transpose() for d in data.Dlist]
    ns["quadHelperThresh"] = np.array(case.thr)
    ns["quadHelperOver"] = np.array(case.over)
    ns["quadHelperUnder"] = np.array(case.under)
    ns["N"] = case.N
    ns["M"] = case.M
    ns["penalizationweights"] = [None] * case.nb      # Laptop variant: Kelly measure per control point
    return ns


def variant_namespace(case, module: str, gastep=None) -> dict:
    """Namespace with the functions of greedyVMAT.py / greedyVMATshort.py bound to `case`: these generations
    read the module globals ``Dlist`` (not data.Dlist), ``quadHelper*``, ``gastep``, ``N``, ``M`` and keep
    data.notinC / data.caligraphicC as plain lists (Appendix A step 3)."""
    ns = extract(module)
    data = ns["vmat_class"]()
    data.numvoxels = case.V
    data.numbeams = case.nb
    data.xinter = np.array(case.xinter)
    data.yinter = np.array(case.yinter)
    data.xdirection = [np.array(a) for a in case.xdirection]
    data.ydirection = [np.array(a) for a in case.ydirection]
    data.llist = [[-1] * case.M for _ in range(case.nb)]
    data.rlist = [[case.N + 1] * case.M for _ in range(case.nb)]
    data.openApertureMaps = [[] for _ in range(case.nb)]
    data.diagmakers = [[] for _ in range(case.nb)]
    data.currentIntensities = np.zeros(case.nb)
    data.notinC = list(range(case.nb))
    data.caligraphicC = []
    ns["data"] = data
    ns["Dlist"] = case.Dlist()
    ns["DlistT"] = [d.transpose() for d in ns["Dlist"]]          # greedyVMAT.py:124 (greedyVMATshort.py transposes in place)
    ns["quadHelperThresh"] = np.array(case.thr)
    ns["quadHelperOver"] = np.array(case.over)
    ns["quadHelperUnder"] = np.array(case.under)
    ns["N"], ns["M"] = case.N, case.M
    ns["gastep"] = gastep if gastep is not None else (case.angles[1] - case.angles[0] if case.nb > 1 else 1)
    return ns


def variant_colgen(ns, case, module: str, C, max_iter=None, C2=1.0, C3=1.0, b=0.5):
    """The colGen loops of greedyVMATshort.py:746-814 / greedyVMAT.py:779-836 (WholeCircle=False) driven over the
    reference's own PricingProblem / updateOpenAperture / solveRMC, without plots and file output."""
    import warnings
    data = ns["data"]
    short = module == "greedyVMATshort"
    vmax, speedlim = (2.0, 3.0) if short else (2.0 * 1000, 3.0)
    data.llist = [[-1] * case.M for _ in range(case.nb)]
    data.rlist = [[case.N + 1] * case.M for _ in range(case.nb)]
    with quiet():
        data.calcDose()
    data.caligraphicC = []
    data.notinC = list(range(case.nb))
    pstar = -float("inf")
    history = []
    while pstar < 0 and (short or len(data.notinC) > 0):
        data.calcDose()
        data.calcGradientandObjValue()
        with quiet():
            if short:
                pstar, lm, rm, best = ns["PricingProblem"](C, C2, C3, b, 60, 60, vmax, speedlim, case.N, case.M)
            else:
                pstar, lm, rm, best = ns["PricingProblem"](C, C2, C3, b, vmax, speedlim, case.N, case.M)
        if pstar >= 0:
            break
        data.caligraphicC.append(best)
        data.caligraphicC.sort()
        data.notinC.remove(best)
        data.notinC.sort()
        data.llist[best], data.rlist[best] = lm, rm
        with quiet():
            data.openApertureMaps[best], data.diagmakers[best] = ns["updateOpenAperture"](best)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = ns["solveRMC"]()
        history.append(dict(idx=int(best), l=[int(v) for v in lm], r=[int(v) for v in rm], pstar=float(pstar),
                            x=np.array(res.x, dtype=float), fun=float(res.fun), nit=int(res.nit)))
        if max_iter is not None and len(history) >= max_iter:
            break
    return history


def ref_pricing_problem(ns, C, C2, C3, b, vmax, speedlim, N, M, bw):
    """Serial stand-in for the reference's PricingProblem (greedyVMATsampling.py:824-856),
    which cannot be extracted because its Pool lives under ``if __name__ == '__main__'``:
    map the reference's own per-candidate function over notinC.loc, np.argmin of p."""
    data = ns["data"]
    with quiet():
        res = [ns["parallelizationPricingProblem"](i, C, C2, C3, b, vmax, speedlim, N, M, bw)
               for i in data.notinC.loc]
    pvalues = np.array([r[0] for r in res])
    k = int(np.argmin(pvalues))
    return res[k][0], res[k][1], res[k][2], res[k][3], res


def ref_colgen(ns, case, C, kappasize, max_iter=None, C2=1.0, C3=1.0, b=0.5, vmax=2.25,
               speedlim=0.83, bw=0.5, RU=10.0, elimination=False, threshold=10E-3):
    """The reference's colGen loop (greedyVMATsampling.py:937-1046) driven over the
    reference's own functions, without its plotting / pickle side effects.
    Returns the list of (bestApertureIndex, l, r, pstar, rmp.x copy, rmp.fun)."""
    data = ns["data"]
    YU = RU / speedlim
    kappa = list(range(case.nb))
    random.seed(13)
    kappa = random.sample(kappa, len(kappa))
    for _ in range(min(kappasize, len(kappa))):
        i = kappa.pop(0)
        data.notinC.insertAngle(i, data.pointtoAngle[i])
    data.caligraphicC = ns["apertureList"]()
    pstar = -np.inf
    history = []
    while (pstar < 0) and (data.notinC.len() > 0):
        data.calcDose()
        data.calcGradientandObjValue()
        with quiet():      # the reference's own PricingProblem, its Pool replaced by a serial map
            pstar, lm, rm, best = ns["PricingProblem"](C, C2, C3, b, vmax, speedlim, case.N, case.M, bw)
        if pstar >= 0:
            break
        data.caligraphicC.insertAngle(best, data.notinC(best))
        data.notinC.removeIndex(best)
        data.llist[best] = lm
        data.rlist[best] = rm
        with quiet():
            oam, dm, st = ns["updateOpenAperture"](best)
        data.openApertureMaps[best], data.diagmakers[best], data.strengths[best] = oam, dm, st
        with quiet():
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = ns["solveRMC"](YU)
        data.rmpres = res
        removed = []                                   # elimination phase (:1007-1033)
        for k in range(case.nb):
            if k in data.caligraphicC.loc and (res.x[k] < threshold) and elimination:
                data.entryCounter += 1
                removed.append(k)
                data.notinC.insertAngle(k, data.pointtoAngle[k])
                data.caligraphicC.removeIndex(k)
        stop = False
        prev = data.listIndexofAperturesRemovedEachStep
        if len(prev) > 1 and (np.any(np.isin(removed, prev[-1])) or np.any(np.isin(removed, prev[-2]))):
            stop = True
        if not stop:
            prev.append(removed)
        history.append(dict(idx=int(best), l=[int(v) for v in lm], r=[int(v) for v in rm],
                            pstar=float(pstar), x=np.array(res.x, dtype=float), fun=float(res.fun),
                            nit=int(res.nit), nfev=int(res.nfev), removed=list(removed)))
        if stop:
            break
        while not data.notinC.isEmpty():
            kappa.append(data.notinC.loc[0])
            data.notinC.removeIndex(data.notinC.loc[0])
        random.seed(13)
        for i in random.sample(kappa, min(kappasize, len(kappa))):
            data.notinC.insertAngle(i, data.pointtoAngle[i])
            kappa.remove(i)
        if max_iter is not None and len(history) >= max_iter:
            break
    return history


def run_reference_loader(cfg: dict, module: str = "greedyVMATsampling", with_functions: bool = False):
    """Execute the reference's own loader - the TOP-LEVEL statements of the script between
    ``start = time.time()`` and the quadHelper loops (greedyVMATsampling.py:286-576), plus the
    class definitions they need - on a data set written by oracle/cort_fixture.py.  The hard-coded
    path / priority / gantry-range assignments are dropped and replaced by `cfg`.
    Returns the namespace (``data``, ``DlistT``, ``quadHelper*``, ``N``, ``M``, ...).  With
    `with_functions` the reference's hot-path functions are defined in the same namespace, so the
    greedy loop can then run on what the reference's loader produced (``ref_colgen_loaded``)."""
    import glob
    import scipy.io as sio
    path = os.path.join(REFERENCE_ROOT, module + ".py")
    tree = ast.parse(open(path).read())
    seeded = {"readfolder", "readfolderD", "outputfolder", "objfile", "structurefile", "algfile", "priority",
              "gastart", "gaend", "gastep", "WholeCircle", "IntheOffice", "AtHome", "kappa"}
    first = next(n.lineno for n in tree.body
                 if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "start")
    last = max(n.end_lineno for n in tree.body
               if isinstance(n, ast.For) and "quadHelperThresh" in ast.dump(n))
    keep = []
    for n in tree.body:
        if isinstance(n, ast.ClassDef) and n.name in ("region", "apertureList", "vmat_class"):
            keep.append(n)
        elif with_functions and isinstance(n, ast.FunctionDef) and n.name in _HOT:
            keep.append(n)
        elif first <= n.lineno <= last:
            if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") in seeded:
                continue
            keep.append(n)
    ns = extract(module, names=[])          # shimmed numpy, SerialPool, __name__ == '__main__'
    ns.update(glob=glob, os=os, sio=sio, ga=[], ca=[], outputfolder=cfg["readfolder"],
              pointtoAngle=range(cfg["gastart"], cfg["gaend"], cfg["gastep"]))
    ns.update({k: (list(v) if k == "priority" else v) for k, v in cfg.items()})
    cwd = os.getcwd()
    try:
        import warnings
        with quiet(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    finally:
        os.chdir(cwd)
    return ns


def ref_colgen_loaded(cfg: dict, C, kappasize, max_iter=None, **kw):
    """Files -> the reference's loader -> the reference's greedy loop: the aperture sequence a user of
    the reference gets from a data set written by oracle/cort_fixture.py."""
    import types
    ns = run_reference_loader(cfg, with_functions=True)
    data = ns["data"]
    M, N = len(data.xinter), len(data.yinter)
    data.llist = [[-1] * M for _ in range(data.numbeams)]
    data.rlist = [[N + 1] * M for _ in range(data.numbeams)]
    data.openApertureMaps = [[] for _ in range(data.numbeams)]
    data.diagmakers = [[] for _ in range(data.numbeams)]
    data.strengths = [[] for _ in range(data.numbeams)]
    data.currentIntensities = np.zeros(data.numbeams)
    data.notinC = ns["apertureList"]()
    data.caligraphicC = ns["apertureList"]()
    data.listIndexofAperturesRemovedEachStep = []
    data.pointtoAngle = list(data.pointtoAngle)
    ns["penalizationweights"] = [None] * data.numbeams
    shape = types.SimpleNamespace(nb=data.numbeams, N=N, M=M)
    return ref_colgen(ns, shape, C, kappasize, max_iter=max_iter, **kw), ns


def ref_printresults(dose, full_mask, numstructs):
    """The reference's own printresults (Laptop variant, greedyVMATsamplingLaptop.py:902-949) with
    plotting modules that only record: returns the (bin_center, dvh_matrix) of its first plot call."""
    import types
    ns = extract("greedyVMATsamplingLaptop", names=["printresults"])
    calls = []
    fake = types.SimpleNamespace(plot=lambda *a, **k: calls.append(a), xlim=lambda *a, **k: None)
    noop = lambda *a, **k: None
    ns["pylab"] = fake
    ns["plt"] = types.SimpleNamespace(grid=noop, xlabel=noop, ylabel=noop, title=noop, legend=noop, savefig=noop, close=noop)
    ns["data"] = types.SimpleNamespace(fullMaskValue=full_mask, currentDose=np.asarray(dose), numstructs=numstructs)
    ns["allNames"] = ["S%02d_VOILIST.mat" % s for s in range(numstructs)]
    try:
        with quiet():
            ns["printresults"](1, "/tmp/", 0)
    except IndexError:          # the second plot picks structures 7..20 of the head-and-neck case
        pass
    bin_center, dvh_t = calls[0][0], calls[0][1]
    return np.asarray(bin_center), np.asarray(dvh_t).T
