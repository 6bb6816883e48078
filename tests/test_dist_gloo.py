"""world_size-2 gloo test of the voxel-slab sharding logic (CPU): each rank evaluates its
slab with a host engine built on the oracle, the [gb | f] buffer is all-reduced, and the
result must equal the unsharded evaluation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class HostSlabEngine:
    """DosePlan-shaped engine over the numpy oracle, for exercising ShardedEvaluator on CPU."""

    def __init__(self, case, rows, D, thr, over, under):
        self.case, self.D, self.thr, self.over, self.under = case, D, thr, over, under
        self.buf = torch.zeros(case.B + 1, dtype=torch.float64)
        self.w_dose = np.zeros(case.B); self.w_grad = np.zeros(case.B)
        self.y = np.zeros(case.nb); self.result = None

    def set_weights(self, wd, wg):
        self.w_dose, self.w_grad = np.array(wd), np.array(wg)

    def set_intensities(self, y):
        self.y = np.array(y)

    def eval_async(self, stream, phase):
        from oracle.vmat_oracle import penalty
        if phase in (0, 1):
            x = np.repeat(self.y, np.diff(self.case.beam_ptr)) * self.w_dose
            d = self.D @ x
            f, vg = penalty(d, self.thr, self.over, self.under)
            self.buf[:-1] = torch.from_numpy(self.D.T @ vg)
            self.buf[-1] = float(f)
        if phase in (0, 2):
            gb = self.buf[:-1].numpy()
            g = np.add.reduceat(self.w_grad * gb, self.case.beam_ptr[:-1].astype(np.int64))
            self.result = (float(self.buf[-1]), g)

    def reduce_buffer(self):
        return self.buf

    def fetch_result(self):
        return self.result


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vmatwpencode_b200.synth import make_case
    from vmatwpencode_b200.dist import ShardedEvaluator, slab_of
    case = make_case(V=5000, nb=8, M=4, N=6, density=0.03, seed=91)
    rows, D, thr, over, under = slab_of(case, rank, world)
    eng = HostSlabEngine(case, rows, D, thr, over, under)
    rng = np.random.default_rng(0)
    wd = (rng.random(case.B) > 0.3) * 1.0
    wg = (rng.random(case.B) > 0.3) * 1.0
    y = rng.random(case.nb) * 2
    eng.set_weights(wd, wg)
    f, g = ShardedEvaluator(eng).eval(y)
    if rank == 0:
        np.savez(out, f=f, g=g, wd=wd, wg=wg, y=y)
    dist.destroy_process_group()


def test_two_rank_sharded_eval_equals_unsharded(tmp_path):
    from oracle import vmat_oracle as O
    from vmatwpencode_b200.synth import make_case
    from vmatwpencode_b200.dist import partition_voxels
    out = str(tmp_path / "r0.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z = np.load(out)
    case = make_case(V=5000, nb=8, M=4, N=6, density=0.03, seed=91)
    f0, g0, *_ = O.calc_obj_grad_clean(O.OracleState(case), z["y"], z["wd"], z["wg"])
    assert abs(float(z["f"]) - f0) <= 1e-12 * abs(f0)
    assert np.abs(z["g"] - g0).max() <= 1e-12 * np.abs(g0).max()
    b = partition_voxels(case.D.indptr, 4)
    assert b[0] == 0 and b[-1] == case.V and all(x <= y for x, y in zip(b, b[1:]))
    nnz = np.diff(case.D.indptr[b])
    assert nnz.max() <= 1.3 * nnz.mean()
