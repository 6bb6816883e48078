"""Two-GPU test of the voxel-sharded evaluation: the fused peer-memory exchange and the NCCL
path must both reproduce the single-GPU result.  Skipped with fewer than two CUDA devices."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from vmatwpencode_b200.synth import make_case
    from vmatwpencode_b200.dist import ShardedEvaluator, slab_of
    from vmatwpencode_b200.plan import DosePlan
    case = make_case(V=50_000, nb=16, M=5, N=8, density=0.01, seed=93)
    rows, D, thr, over, under = slab_of(case, rank, world)
    rng = np.random.default_rng(0)
    wd = (rng.random(case.B) > 0.3) * 1.0
    wg = (rng.random(case.B) > 0.3) * 1.0
    ys = [rng.random(case.nb) * 2 for _ in range(3)]
    res = {}
    for mode in ("fused", "nccl"):
        plan = DosePlan(D, case.beam_ptr, thr, over, under, device=rank, store="f32", acc="f64")
        plan.set_weights(wd, wg)
        ev = ShardedEvaluator(plan, fused=(mode == "fused"))
        assert ev.fused == (mode == "fused")
        side = torch.cuda.Stream()
        torch.cuda.set_stream(side)
        vals = []
        for k in range(6):                       # repeated evaluations exercise both buffer parities
            f, g = ev.eval(ys[k % 3], stream=side.cuda_stream)
            vals.append(np.concatenate([[f], g]))
        gb = plan.beamlet_grad()
        res[mode] = (np.array(vals), gb)
        if mode == "fused":                      # host-buffer entry point on the fused path
            f, g = plan.eval(ys[1])
            assert np.array_equal(np.concatenate([[f], g]), vals[1])
        dist.barrier()
        plan.close()
    if rank == 0:
        np.savez(out, fused=res["fused"][0], nccl=res["nccl"][0], gb_fused=res["fused"][1], gb_nccl=res["nccl"][1],
                 wd=wd, wg=wg, ys=np.array(ys))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_evaluation(tmp_path):
    import torch.multiprocessing as mp
    from oracle import vmat_oracle as O
    from vmatwpencode_b200.synth import make_case
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, 29700 + os.getpid() % 200, out), nprocs=2, join=True)
    z = np.load(out)
    case = make_case(V=50_000, nb=16, M=5, N=8, density=0.01, seed=93)
    st = O.OracleState(case)
    for k in range(6):
        f0, g0, d0, vg0, gb0 = O.calc_obj_grad_clean(st, z["ys"][k % 3], z["wd"], z["wg"])
        want = np.concatenate([[f0], g0])
        for mode in ("fused", "nccl"):
            got = z[mode][k]
            assert np.abs(got - want).max() <= 1e-11 * np.abs(want).max(), (mode, k)
    assert np.abs(z["gb_fused"] - gb0).max() <= 1e-11 * np.abs(gb0).max()
    # the fused exchange adds the ranks in rank order on every rank: evaluations repeat bit for bit
    assert np.array_equal(z["fused"][0], z["fused"][3]) and np.array_equal(z["fused"][1], z["fused"][4])


def _greedy_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from vmatwpencode_b200 import VMATlibrary as vl
    from vmatwpencode_b200.synth import make_case
    case = make_case(V=20_000, nb=16, M=5, N=8, density=0.01, seed=94, gastep=10)
    data = vl.bind(case, device=rank, store="f32", acc="f64", shard=True)
    vl.colGen(0.01, False, 6, max_iter=8)
    hist = [(h["idx"], h["l"], h["r"], h["fun"]) for h in data.history]
    gathered = [None] * world
    dist.all_gather_object(gathered, hist)
    if rank == 0:
        np.save(out, np.array([g == gathered[0] for g in gathered]))
        import pickle
        with open(out + ".pkl", "wb") as f:
            pickle.dump(hist, f)
    dist.barrier()
    data.plan.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_greedy_loop_matches_oracle(tmp_path):
    """The whole greedy loop on a voxel-sharded plan: both ranks take the same decisions (the
    fused exchange gives them identical bits) and the aperture sequence is the oracle's."""
    import pickle
    import torch.multiprocessing as mp
    from oracle import vmat_oracle as O
    from vmatwpencode_b200.synth import make_case
    out = str(tmp_path / "same.npy")
    mp.spawn(_greedy_worker, args=(2, 29900 + os.getpid() % 90, out), nprocs=2, join=True)
    assert np.load(out).all()
    hist = pickle.load(open(out + ".pkl", "rb"))
    case = make_case(V=20_000, nb=16, M=5, N=8, density=0.01, seed=94, gastep=10)
    want = O.col_gen(O.OracleState(case), C=0.01, kappasize=6, max_iter=8, clean=True)
    assert [h[0] for h in hist] == [w["idx"] for w in want] and len(want) > 0
    for h, w in zip(hist, want):
        assert h[1] == w["l"] and h[2] == w["r"]


def _km_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from vmatwpencode_b200 import kmedoids as km
    rng = np.random.default_rng(11)
    P = np.round(rng.random((3000, 3)) * 64)            # voxel-like integer coordinates
    got = km.givemewholelist([0, 7], P, device=rank, group=dist.group.WORLD, limit=12)
    if rank == 0:
        np.save(out, np.array(got))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_kmedoid_voxel_sampling(tmp_path):
    import torch.multiprocessing as mp
    from vmatwpencode_b200 import kmedoids as km
    out = str(tmp_path / "km.npy")
    mp.spawn(_km_worker, args=(2, 29990 + os.getpid() % 9, out), nprocs=2, join=True)
    rng = np.random.default_rng(11)
    P = np.round(rng.random((3000, 3)) * 64)
    single = km.givemewholelist([0, 7], P, limit=12)
    assert list(np.load(out)) == single


def test_fused_exchange_two_ranks_on_one_device():
    """The push exchange of the reduction kernel with both "ranks" living on one GPU (two plans of this
    process attached to each other through vmat_comm_attach_local), so that the fused path is exercised
    on a single-GPU box too.  The two evaluations are enqueued on the plans' own streams before either
    result is fetched; the plans are kept small (short rings) so that the kernels of both can be resident
    together - rank 0's reduction kernel spins until rank 1's has delivered."""
    from oracle import vmat_oracle as O
    from vmatwpencode_b200.synth import make_case
    from vmatwpencode_b200.dist import slab_of
    from vmatwpencode_b200.plan import DosePlan
    from vmatwpencode_b200._lib import VmatError
    case = make_case(V=50_000, nb=16, M=5, N=8, density=0.01, seed=93)
    rng = np.random.default_rng(0)
    wd = (rng.random(case.B) > 0.3) * 1.0
    wg = (rng.random(case.B) > 0.3) * 1.0
    ys = [rng.random(case.nb) * 2 for _ in range(3)]
    st = O.OracleState(case)
    # three-kernel path (the exchange sits in the small reduction kernel) and the fused evaluation kernel with
    # grids of 60 CTAs, so that both "ranks" fit on the 148 SMs at once
    for acc, tol, opts in (("f64", 1e-11, dict(fused=1, k1_stages=3, k2_stages=4, k2_tasks=64)),
                           ("f32", 1e-5, dict(fused=1, k1_stages=3, k2_stages=4, k2_tasks=64)),
                           ("f64", 1e-11, dict(fused=2, k1_stages=4, k2_tasks=60)),
                           ("f32", 1e-5, dict(fused=2, k1_stages=4, k2_tasks=60))):
        plans = []
        for rank in range(2):
            rows, D, thr, over, under = slab_of(case, rank, 2)
            p = DosePlan(D, case.beam_ptr, thr, over, under, device=0, store="f32", acc=acc, **opts)
            p.set_weights(wd, wg)
            plans.append(p)
        for rank in range(2):
            plans[rank].comm_attach_local(rank, plans)
            plans[rank].comm_set_timeout(5.0)
        seen = []
        for k in range(6):                                   # both buffer parities, three times each
            for p in plans:
                p.set_intensities(ys[k % 3])
            for p in plans:
                p.eval_async(None, 0)
            res = [np.concatenate([[f], g]) for f, g in (p.fetch_result() for p in plans)]
            assert np.array_equal(res[0], res[1])            # rank order sums: identical bits on both ranks
            f0, g0, d0, vg0, gb0 = O.calc_obj_grad_clean(st, ys[k % 3], wd, wg)
            want = np.concatenate([[f0], g0])
            assert np.abs(res[0] - want).max() <= tol * np.abs(want).max(), (acc, k)
            seen.append(res[0])
            for p in plans:
                gb = p.beamlet_grad()
                assert np.abs(gb - gb0).max() <= tol * np.abs(gb0).max()
        assert np.array_equal(seen[0], seen[3]) and np.array_equal(seen[1], seen[4])
        # a rank that evaluates alone is reported, not hung and not silently wrong
        plans[0].comm_set_timeout(0.2)
        with pytest.raises(VmatError, match="timed out"):
            plans[0].eval(ys[0])
        # ... and after attaching again the pair works as before
        for rank in range(2):
            plans[rank].comm_attach_local(rank, plans)
            plans[rank].comm_set_timeout(5.0)
        for p in plans:
            p.set_intensities(ys[1])
        for p in plans:
            p.eval_async(None, 0)
        res = [np.concatenate([[f], g]) for f, g in (p.fetch_result() for p in plans)]
        assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], seen[1])
        for p in plans:
            p.close()
