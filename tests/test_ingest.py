"""CORT-format ingest (vmatwpencode_b200/ingest.py) against what the reference's own loader
statements produced for the same synthetic data set (fixtures: oracle/make_golden.py
ingest_fixture; live when /root/reference is mounted).  CPU only, plus one GPU test that binds
the loaded case."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.cort_fixture import write_cort_like
from oracle import ref_extract as R
from vmatwpencode_b200.ingest import load_cort
from helpers import GOLDEN


def _load(tmp_path, seed):
    cfg = write_cort_like(str(tmp_path / f"cort{seed}"), seed=seed)
    case = load_cort(cfg["readfolder"], cfg["readfolderD"], cfg["structurefile"], cfg["objfile"], cfg["priority"],
                     cfg["gastart"], cfg["gaend"], cfg["gastep"], cfg["algfile"])
    return cfg, case


def _check(case, z):
    assert case.V == int(z["numvoxels"]) and case.nb == int(z["numbeams"])
    assert case.N == int(z["N"]) and case.M == int(z["M"])
    assert case.numX == int(z["numX"]) and case.totaldijs == int(z["totaldijs"])
    assert case.gantry == list(z["ga"]) and case.angles == list(z["pointtoAngle"])
    assert np.array_equal(case.xinter, z["xinter"]) and np.array_equal(case.yinter, z["yinter"])
    assert np.array_equal(np.diff(case.beam_ptr), z["beamletsPerBeam"])
    assert np.array_equal(np.concatenate(case.xdirection), z["xdir"])
    assert np.array_equal(np.concatenate(case.ydirection), z["ydir"])
    assert np.array_equal(case.voxelAssignment, z["voxelAssignment"])
    assert np.array_equal(case.originalVoxels, z["originalVoxels"], equal_nan=True)
    assert np.array_equal(case.maskValue, z["maskValue"]) and np.array_equal(case.fullMaskValue, z["fullMaskValue"])
    assert [r.index for r in case.regions] == list(z["region_index"])
    assert [bool(r.target) for r in case.regions] == [bool(t) for t in z["region_target"]]
    assert [r.sizeInVoxels for r in case.regions] == list(z["region_size"])
    assert np.array_equal(np.concatenate([r.indices for r in case.regions]), z["region_indices"])
    assert np.array_equal(np.concatenate([np.ravel(r.fullIndices) for r in case.regions]), z["region_full"])
    assert np.array_equal(case.functionData, z["functionData"])
    for a, b in ((case.thr, z["thr"]), (case.over, z["over"]), (case.under, z["under"])):
        assert np.array_equal(a, b)
    D = case.D.tocsr(); D.sort_indices()
    assert np.array_equal(D.indptr, z["D_indptr"]) and np.array_equal(D.indices, z["D_indices"])
    assert np.array_equal(D.data, z["D_data"])


@pytest.mark.parametrize("name", ["ingest_A", "ingest_B"])
def test_ingest_matches_reference_fixture(tmp_path, name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg, case = _load(tmp_path, int(z["seed"]))
    _check(case, z)


@pytest.mark.skipif(not R.reference_available(), reason="reference not mounted")
def test_ingest_matches_reference_live(tmp_path):
    cfg, case = _load(tmp_path, 7)
    ns = R.run_reference_loader(cfg)
    data = ns["data"]
    assert case.V == data.numvoxels and case.nb == data.numbeams
    assert np.array_equal(case.thr, ns["quadHelperThresh"]) and np.array_equal(case.over, ns["quadHelperOver"])
    assert np.array_equal(case.under, ns["quadHelperUnder"])
    for i, Dl in enumerate(case.Dlist()):
        ref = sp.csr_matrix(data.Dlist[i]); ref.sort_indices(); Dl.sort_indices()
        assert ref.shape == Dl.shape and np.array_equal(ref.indptr, Dl.indptr)
        assert np.array_equal(ref.indices, Dl.indices) and np.array_equal(ref.data, Dl.data)
    assert np.array_equal(case.fullMaskValue, data.fullMaskValue)


def test_files_to_aperture_sequence_oracle_matches_reference_fixture(tmp_path):
    """Data set on disk -> load_cort -> the oracle's greedy loop, against what the reference's own
    loader and loop produced from the same files (fixture colgen_files): bit for bit."""
    from oracle import vmat_oracle as O
    z = np.load(os.path.join(GOLDEN, "colgen_files.npz"))
    cfg, case = _load(tmp_path, int(z["seed"]))
    hist = O.col_gen(O.OracleState(case), C=float(z["C"]), kappasize=int(z["kappasize"]), max_iter=int(z["max_iter"]))
    assert [h["idx"] for h in hist] == list(z["idx"]) and len(hist) > 0
    for h, l, r, p, x, fun in zip(hist, z["l"], z["r"], z["pstar"], z["x"], z["fun"]):
        assert h["l"] == list(l) and h["r"] == list(r)
        assert h["pstar"] == p and h["fun"] == fun and np.array_equal(h["x"], x)


@pytest.mark.skipif(not R.reference_available(), reason="reference not mounted")
def test_files_to_aperture_sequence_live(tmp_path):
    from oracle import vmat_oracle as O
    cfg, case = _load(tmp_path, 5)
    ref, _ = R.ref_colgen_loaded(cfg, 0.01, 3, max_iter=5)
    hist = O.col_gen(O.OracleState(case), C=0.01, kappasize=3, max_iter=5)
    assert [(h["idx"], h["l"], h["r"], h["pstar"], h["fun"]) for h in hist] == \
           [(h["idx"], h["l"], h["r"], h["pstar"], h["fun"]) for h in ref]


@pytest.mark.gpu
def test_loaded_case_runs_the_greedy_loop(tmp_path):
    """Data set on disk -> load_cort -> bind -> colGen on the GPU: the aperture sequence the
    reference's own loader and loop produced from the same files (fixture colgen_files; ragged
    beamlet rows come from the grid cut here, not from the generator)."""
    from vmatwpencode_b200 import VMATlibrary as vl
    z = np.load(os.path.join(GOLDEN, "colgen_files.npz"))
    cfg, case = _load(tmp_path, int(z["seed"]))
    data = vl.bind(case, store="f64", acc="f64")
    vl.colGen(float(z["C"]), False, int(z["kappasize"]), max_iter=int(z["max_iter"]))
    hist = data.history
    assert [h["idx"] for h in hist] == list(z["idx"]) and len(hist) > 0
    for h, l, r, p, fun in zip(hist, z["l"], z["r"], z["pstar"], z["fun"]):
        assert h["l"] == list(l) and h["r"] == list(r)          # bit-exact leaf positions
        assert abs(h["pstar"] - p) <= 1e-9 * abs(p) and abs(h["fun"] - fun) <= 1e-9 * abs(fun)
    data.plan.close()
