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


@pytest.mark.gpu
def test_loaded_case_runs_the_greedy_loop(tmp_path):
    """A loaded case binds like a synthetic one and the greedy loop matches the oracle on it
    (ragged beamlet rows come from the grid cut here, not from the generator)."""
    from oracle import vmat_oracle as O
    from vmatwpencode_b200 import VMATlibrary as vl
    cfg, case = _load(tmp_path, 3)
    want = O.col_gen(O.OracleState(case), C=0.01, kappasize=3, max_iter=4, clean=True)
    data = vl.bind(case, store="f64", acc="f64")
    vl.colGen(0.01, False, 3, max_iter=4)
    assert [h["idx"] for h in data.history] == [h["idx"] for h in want] and len(want) > 0
    for a, b in zip(data.history, want):
        assert a["l"] == b["l"] and a["r"] == b["r"]
    data.plan.close()
