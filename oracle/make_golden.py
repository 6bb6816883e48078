"""TEST INFRASTRUCTURE - freeze outputs of the reference's own functions into tests/golden/.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
Each fixture stores the synthetic inputs together with what the reference's AST-extracted
functions (oracle/ref_extract.py) returned for them, so the pin travels to machines without
the reference.  Nothing here copies reference code; only its outputs are recorded.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_extract as R                                   # noqa: E402
from vmatwpencode_b200.synth import make_case, case_to_arrays          # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def select(ns, case, beam, l, r):
    """Select control point `beam` with leaf vectors l, r through the reference's own
    updateOpenAperture."""
    data = ns["data"]
    if beam in data.notinC.loc:
        data.notinC.removeIndex(beam)
    data.caligraphicC.insertAngle(beam, case.angles[beam])
    data.llist[beam] = list(l)
    data.rlist[beam] = list(r)
    with R.quiet():
        out = ns["updateOpenAperture"](beam)
    data.openApertureMaps[beam], data.diagmakers[beam], data.strengths[beam] = out
    return out


def eval_and_price_fixture(name, case, selected, y, price_settings):
    ns = R.reference_namespace(case)
    data = ns["data"]
    for i in range(case.nb):
        data.notinC.insertAngle(i, case.angles[i])
    out = case_to_arrays(case)
    sel_beams = []
    for beam, (l, r) in selected.items():
        oam, dm, stg = select(ns, case, beam, l, r)
        sel_beams.append(beam)
        out[f"sel{beam}_l"] = np.asarray(l, dtype=float)
        out[f"sel{beam}_r"] = np.asarray(r, dtype=float)
        out[f"sel{beam}_oam"] = np.asarray(oam, dtype=np.int64)
        out[f"sel{beam}_diag"] = np.asarray(dm, dtype=float)
        out[f"sel{beam}_strength"] = np.asarray(stg, dtype=float)
    out["selected"] = np.asarray(sel_beams, dtype=np.int64)
    out["y"] = np.asarray(y, dtype=float)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f, g = ns["calcObjGrad"](np.array(y, dtype=float))
    out["f"] = np.float64(f)
    out["g"] = np.asarray(g, dtype=float).ravel()
    out["dose"] = np.asarray(data.currentDose, dtype=float)
    out["vg"] = np.asarray(data.voxelgradient, dtype=float)
    settings = []
    for k, (C, vmax, speedlim, bw) in enumerate(price_settings):
        ps, ls, rs, idxs = [], [], [], []
        for i in list(data.notinC.loc):
            with R.quiet():
                p, l, r, idx = ns["parallelizationPricingProblem"](i, C, 1.0, 1.0, 0.5, vmax, speedlim, case.N, case.M, bw)
            ps.append(float(p)); ls.append([int(v) for v in l]); rs.append([int(v) for v in r]); idxs.append(idx)
        out[f"price{k}_params"] = np.array([C, 1.0, 1.0, 0.5, vmax, speedlim, bw])
        out[f"price{k}_idx"] = np.asarray(idxs, dtype=np.int64)
        out[f"price{k}_p"] = np.asarray(ps)
        out[f"price{k}_l"] = np.asarray(ls, dtype=np.int64)
        out[f"price{k}_r"] = np.asarray(rs, dtype=np.int64)
        settings.append(k)
    out["n_price"] = np.int64(len(settings))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "f =", float(f), "candidates", len(data.notinC.loc))


def colgen_fixture(name, case, C, kappasize, max_iter, elimination=False, module="greedyVMATsampling"):
    ns = R.reference_namespace(case, module)
    hist = R.ref_colgen(ns, case, C=C, kappasize=kappasize, max_iter=max_iter, elimination=elimination)
    out = case_to_arrays(case)
    out["C"] = np.float64(C)
    out["kappasize"] = np.int64(kappasize)
    out["max_iter"] = np.int64(max_iter)
    out["idx"] = np.asarray([h["idx"] for h in hist], dtype=np.int64)
    out["l"] = np.asarray([h["l"] for h in hist], dtype=np.int64)
    out["r"] = np.asarray([h["r"] for h in hist], dtype=np.int64)
    out["pstar"] = np.asarray([h["pstar"] for h in hist])
    out["x"] = np.asarray([h["x"] for h in hist])
    out["fun"] = np.asarray([h["fun"] for h in hist])
    out["elimination"] = np.bool_(elimination)
    out["removed_flat"] = np.asarray([k for h in hist for k in h["removed"]], dtype=np.int64)
    out["removed_count"] = np.asarray([len(h["removed"]) for h in hist], dtype=np.int64)
    if module.endswith("Laptop"):      # Kelly measure the reference stored for the control points still selected
        pw = ns["penalizationweights"]
        out["kelly_idx"] = np.asarray([k for k, v in enumerate(pw) if v is not None], dtype=np.int64)
        out["kelly"] = np.asarray([float(v) for v in pw if v is not None])
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "iterations", len(hist), "sequence", [h["idx"] for h in hist])


def variant_fixture(name, module, case, C, max_iter, gastep):
    """A greedy run of one of the EARLIER generations (greedyVMATshort.py / greedyVMAT.py) by the reference's own
    functions, plus its per-candidate pricing results in the state the run ends in."""
    ns = R.variant_namespace(case, module, gastep=gastep)
    hist = R.variant_colgen(ns, case, module, C=C, max_iter=max_iter)
    data = ns["data"]
    out = case_to_arrays(case)
    out["C"] = np.float64(C)
    out["max_iter"] = np.int64(max_iter)
    out["gastep"] = np.int64(gastep)
    out["idx"] = np.asarray([h["idx"] for h in hist], dtype=np.int64)
    out["l_flat"] = np.asarray([v for h in hist for v in h["l"]], dtype=np.int64)
    out["r_flat"] = np.asarray([v for h in hist for v in h["r"]], dtype=np.int64)
    out["l_len"] = np.asarray([len(h["l"]) for h in hist], dtype=np.int64)
    out["pstar"] = np.asarray([h["pstar"] for h in hist])
    out["x"] = np.asarray([h["x"] for h in hist])
    out["fun"] = np.asarray([h["fun"] for h in hist])
    # pricing of every remaining candidate in the final state (the evaluation of that state first)
    data.calcDose()
    data.calcGradientandObjValue()
    ps, ls, rs, idxs = [], [], [], []
    short = module == "greedyVMATshort"
    for k, i in enumerate(list(data.notinC)):
        with R.quiet():
            if short:
                p, l, r = ns["PPsubroutine"](C, 1.0, 1.0, 0.5, 60, 60, 2.0, 3.0, [], [], case.N, case.M, i)
            else:
                p, l, r, _ = ns["parallelizationPricingProblem"](k, C, 1.0, 1.0, 0.5, 2.0 * 1000, 3.0, case.N, case.M)
        ps.append(float(p)); ls.append([int(v) for v in l]); rs.append([int(v) for v in r]); idxs.append(int(i))
    out["final_idx"] = np.asarray(idxs, dtype=np.int64)
    out["final_p"] = np.asarray(ps)
    out["final_l_flat"] = np.asarray([v for l in ls for v in l], dtype=np.int64)
    out["final_r_flat"] = np.asarray([v for r in rs for v in r], dtype=np.int64)
    out["final_len"] = np.asarray([len(l) for l in ls], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "iterations", len(hist), "sequence", [h["idx"] for h in hist], "remaining", idxs)


def variant_fixtures():
    """BASELINE.json configs[0] names greedyVMATshort.py; greedyVMAT.py is the generation in between."""
    variant_fixture("variant_short_A", "greedyVMATshort", make_case(V=1500, nb=6, M=5, N=6, density=0.05, seed=11, gastep=60), 0.5, 4, 60)
    variant_fixture("variant_short_B", "greedyVMATshort", make_case(V=2500, nb=8, M=6, N=8, density=0.03, seed=12, gastep=10, ragged=True), 0.5, 5, 10)
    variant_fixture("variant_classic_A", "greedyVMAT", make_case(V=1500, nb=6, M=5, N=6, density=0.05, seed=11, gastep=60), 0.01, 4, 60)
    variant_fixture("variant_classic_B", "greedyVMAT", make_case(V=2500, nb=8, M=6, N=8, density=0.03, seed=12, gastep=10, ragged=True), 0.5, 5, 10)
    variant_fixture("variant_classic_C", "greedyVMAT", make_case(V=2500, nb=8, M=6, N=8, density=0.03, seed=13, gastep=2), 0.05, 5, 2)


def colgen_files_fixture(name, seed, C, kappasize, max_iter):
    """Files -> the reference's own loader -> the reference's own greedy loop, on a synthetic
    CORT-format data set the tests regenerate from `seed` (oracle/cort_fixture.py)."""
    import tempfile
    from oracle.cort_fixture import write_cort_like
    cfg = write_cort_like(tempfile.mkdtemp(), seed=seed)
    hist, _ = R.ref_colgen_loaded(cfg, C, kappasize, max_iter=max_iter)
    out = dict(seed=np.int64(seed), C=np.float64(C), kappasize=np.int64(kappasize), max_iter=np.int64(max_iter),
               idx=np.asarray([h["idx"] for h in hist], dtype=np.int64),
               l=np.asarray([h["l"] for h in hist], dtype=np.int64), r=np.asarray([h["r"] for h in hist], dtype=np.int64),
               pstar=np.asarray([h["pstar"] for h in hist]), x=np.asarray([h["x"] for h in hist]),
               fun=np.asarray([h["fun"] for h in hist]))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "iterations", len(hist), "sequence", [h["idx"] for h in hist])


def kmedoids_fixture():
    """The reference's only embedded known answer: the hard-coded kappa list
    (greedyVMAT.py:28) equals givemewholelist(range(6,178,11), range(178)); plus small runs
    of the reference's own functions."""
    tree = ast.parse(open(os.path.join(R.REFERENCE_ROOT, "greedyVMAT.py")).read())
    kappa = None
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", None) == "kappa":
            kappa = ast.literal_eval(node.value)
            break
    assert kappa is not None and len(kappa) == 178
    ns = R.extract("kmedoids")
    out = {"kappa178": np.asarray(kappa, dtype=np.int64)}
    for n, start in ((23, [3, 11, 19]), (40, [5, 20, 35]), (31, [0])):
        with R.quiet():
            order = ns["givemewholelist"](list(start), range(0, n))
        out[f"order_n{n}"] = np.asarray(order, dtype=np.int64)
        out[f"start_n{n}"] = np.asarray(start, dtype=np.int64)
        with R.quiet():
            out[f"dmin_n{n}"] = np.float64(ns["distancemin"](list(start), range(0, n)))
    np.savez_compressed(os.path.join(GOLDEN, "kmedoids.npz"), **out)
    print("kmedoids fixtures written")


def kmedoids_perm_fixture():
    """The reference's functions on 1-D point lists that are NOT 0..n-1 in order: its candidates are the
    values of `ba`, visited in `ba`'s order, and double as row indices (kmedoids.py:50-53)."""
    ns = R.extract("kmedoids")
    out = {}
    rng = np.random.default_rng(21)
    for n, start in ((19, [2, 9]), (26, [0])):
        ba = [int(v) for v in rng.permutation(n)]
        with R.quiet():
            order = ns["givemewholelist"](list(start), ba)
        out[f"ba_n{n}"] = np.asarray(ba, dtype=np.int64)
        out[f"start_n{n}"] = np.asarray(start, dtype=np.int64)
        out[f"order_n{n}"] = np.asarray(order, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLDEN, "kmedoids_perm.npz"), **out)
    print("kmedoids permutation fixtures written")


def ingest_fixture(name, seed):
    """What the reference's own loader statements (greedyVMATsampling.py:286-576) produce for a
    synthetic CORT-format data set (oracle/cort_fixture.py, regenerated from `seed` by the tests)."""
    import tempfile
    from oracle.cort_fixture import write_cort_like
    root = tempfile.mkdtemp()
    cfg = write_cort_like(root, seed=seed)
    ns = R.run_reference_loader(cfg)
    data = ns["data"]
    out = dict(seed=np.int64(seed), numvoxels=np.int64(data.numvoxels), numbeams=np.int64(data.numbeams),
               N=np.int64(ns["N"]), M=np.int64(ns["M"]), numX=np.int64(data.numX), totaldijs=np.int64(data.totaldijs),
               ga=np.asarray(ns["ga"], dtype=np.int64), xinter=np.asarray(data.xinter), yinter=np.asarray(data.yinter),
               beamletsPerBeam=np.asarray(data.beamletsPerBeam, dtype=np.int64),
               xdir=np.concatenate(data.xdirection), ydir=np.concatenate(data.ydirection),
               voxelAssignment=np.asarray(data.voxelAssignment), originalVoxels=np.asarray(ns["originalVoxels"]),
               maskValue=np.asarray(data.maskValue), fullMaskValue=np.asarray(data.fullMaskValue),
               thr=ns["quadHelperThresh"], over=ns["quadHelperOver"], under=ns["quadHelperUnder"],
               functionData=np.asarray(data.functionData), pointtoAngle=np.asarray(list(data.pointtoAngle), dtype=np.int64),
               region_index=np.asarray([r.index for r in data.regions], dtype=np.int64),
               region_target=np.asarray([r.target for r in data.regions]),
               region_size=np.asarray([r.sizeInVoxels for r in data.regions], dtype=np.int64),
               region_indices=np.concatenate([np.asarray(r.indices, dtype=np.int64).ravel() for r in data.regions]),
               region_full_size=np.asarray([np.asarray(r.fullIndices).size for r in data.regions], dtype=np.int64),
               region_full=np.concatenate([np.asarray(r.fullIndices, dtype=np.int64).ravel() for r in data.regions]))
    import scipy.sparse as sp
    D = sp.hstack([d.transpose() for d in data.Dlist], format="csr")       # V x B like the kernels take it
    D.sort_indices()
    out.update(D_indptr=D.indptr.astype(np.int64), D_indices=D.indices.astype(np.int32), D_data=D.data)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "voxels", data.numvoxels, "beams", data.numbeams, "grid", ns["M"], "x", ns["N"], "nnz", D.nnz)


def dvh_fixture():
    """Dose-volume histograms from the reference's own printresults (Laptop variant) for a seeded
    dose vector and structure bit masks."""
    rng = np.random.default_rng(17)
    V, nstructs = 3000, 7
    mask = rng.integers(1, 2 ** nstructs, size=V)
    mask[rng.random(V) < 0.3] &= 0b0111011            # structure 2 and 6 thinner; structure 5 never alone
    mask[mask == 0] = 1
    mask &= ~(1 << 5)                                  # structure 5 is empty
    dose = np.round(rng.random(V) * 4.0, 3)            # many values on 0.1 bin edges x.y00
    bin_center, dvh = R.ref_printresults(dose, mask.astype(float), nstructs)
    np.savez_compressed(os.path.join(GOLDEN, "dvh.npz"), dose=dose, full_mask=mask.astype(np.int64),
                        numstructs=np.int64(nstructs), bin_center=bin_center, dvh_matrix=dvh)
    print("dvh fixture", dvh.shape)


def main():
    if not R.reference_available():
        raise SystemExit("reference not mounted at " + R.REFERENCE_ROOT)
    os.makedirs(GOLDEN, exist_ok=True)
    # A: full rectangle grid, two selected control points, loose and tight leaf speeds.
    caseA = make_case(V=1500, nb=10, M=4, N=6, density=0.05, seed=21, gastep=10)
    yA = np.zeros(10); yA[3] = 1.7; yA[6] = 0.9
    eval_and_price_fixture("evalprice_A", caseA, {3: ([0, 0, 1, 1], [5, 6, 5, 4]), 6: ([1, 0, 0, 2], [4, 5, 6, 6])},
                           yA, [(0.0, 2.25, 0.83, 0.5), (0.02, 2.25, 0.83, 0.5), (1.0, 2.25, 0.83, 0.5),
                                (0.02, 0.05, 0.83, 0.5), (0.02, 0.02, 0.83, 0.5)])
    # B: ragged beamlet rows (min(validbeamlets) > 0 in places), gastep 2, three selected.
    caseB = make_case(V=1800, nb=12, M=5, N=7, density=0.05, seed=22, gastep=2, ragged=True)
    yB = np.zeros(12); yB[2] = 2.0; yB[5] = 0.4; yB[9] = 1.1
    eval_and_price_fixture("evalprice_B", caseB,
                           {2: ([0, 1, 1, 0, 2], [6, 5, 7, 4, 5]), 5: ([1, 1, 1, 1, 1], [5, 5, 5, 5, 5]),
                            9: ([-1, 0, 2, 1, 0], [7, 3, 6, 7, 2])},
                           yB, [(0.0, 2.25, 0.83, 0.5), (0.05, 2.25, 0.83, 0.5), (0.05, 0.3, 0.83, 0.5),
                                (0.5, 0.1, 0.83, 0.5)])
    # C: fractional leaves through updateOpenAperture (strength edge cases Q2/Q3).
    caseC = make_case(V=1200, nb=6, M=3, N=8, density=0.06, seed=23, gastep=10)
    yC = np.zeros(6); yC[1] = 1.0; yC[4] = 2.5
    eval_and_price_fixture("evalprice_C", caseC, {1: ([0.5, 2, 2], [5.25, 5, 4]), 4: ([1, -1, 3.75], [6, 8, 5.5])},
                           yC, [(0.01, 2.25, 0.83, 0.5)])
    # D: ragged rows + left leaves left of a row's first beamlet: the open map restarts at index 0
    # and lists beamlets twice (duplicates are deposited twice by calcDose).
    caseD = make_case(V=2000, nb=6, M=6, N=12, density=0.05, seed=80, gastep=2, ragged=True)
    rngD = np.random.default_rng(0)
    selD = {}
    for bm in (1, 3, 4):
        l = [int(rngD.integers(-1, caseD.N - 2)) for _ in range(caseD.M)]
        selD[bm] = (l, [int(min(caseD.N, lv + rngD.integers(1, 6))) for lv in l])
    yD = np.zeros(6); yD[1] = 1.2; yD[3] = 0.8; yD[4] = 1.9
    eval_and_price_fixture("evalprice_D", caseD, selD, yD, [(0.0, 2.25, 0.83, 0.5), (1.0, 2.25, 0.83, 0.5),
                                                            (0.05, 0.1, 0.83, 0.5)])
    colgen_fixture("colgen_A", make_case(V=2500, nb=12, M=5, N=6, density=0.04, seed=31, gastep=10), 0.02, 5, 10)
    colgen_fixture("colgen_B", make_case(V=2000, nb=18, M=4, N=7, density=0.04, seed=32, gastep=2, ragged=True), 0.0, 6, 12)
    colgen_fixture("colgen_C", make_case(V=2000, nb=36, M=3, N=5, density=0.05, seed=33, gastep=1), 0.5, 8, 14)
    colgen_fixture("colgen_E", make_case(V=2500, nb=14, M=4, N=6, density=0.04, seed=35, gastep=10), 0.01, 5, 14,
                   elimination=True, module="greedyVMATsamplingLaptop")
    dvh_fixture()
    ingest_fixture("ingest_A", seed=1)
    ingest_fixture("ingest_B", seed=2)
    colgen_files_fixture("colgen_files", seed=3, C=0.01, kappasize=3, max_iter=6)
    kmedoids_fixture()
    kmedoids_perm_fixture()
    variant_fixtures()


if __name__ == "__main__":
    main()
