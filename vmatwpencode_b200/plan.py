"""DosePlan: host-side handle on one GPU's copy of the dose-deposition matrix.

Thin Python over the C ABI (include/vmat_b200.h).  It replaces what the reference
keeps in ``data.Dlist`` / ``DlistT`` / ``quadHelper*`` (greedyVMATsampling.py:482-483,
524-525,564-576) and what ``calcDose`` / ``calcGradientandObjValue`` / ``PPsubroutine``
compute with scipy.sparse (:243-272,646): one D (V x B) resident in HBM in two
kernel-friendly layouts, and the evaluation  y -> (f, g)  as three CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Options, PriceParams, PriceRow, VmatError, check


def _ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())          # torch tensor


class _DeviceArray:
    """Minimal __cuda_array_interface__ carrier so torch can view a library-owned buffer."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


class DosePlan:
    """One rank's plan: D (its voxel slab) x all beamlets, penalty tables, aperture weights."""

    def __init__(self, D, beam_ptr, thr, over, under, device: int = 0, store: str = "f32", acc: str = "f32",
                 tile_voxels: int = 0, k1_threads: int = 0, k2_threads: int = 0, k2_tasks: int = 0,
                 use_graph: bool = True, Dt=None, chunk_cap: int = 0, pdl: bool = True, host_path: int = 0,
                 index_bytes: int = 0, k1_stages: int = 0, k2_stages: int = 0, claim_lead: int = 0,
                 k2_chunk_cap: int = 0, fused: int = 0, k2_pool_pct: int = 0, consume_batch: int = 0):
        """`D` is a scipy.sparse matrix (V x B, any format) or a dict of CUDA torch tensors
        {vrowptr, vcol, vval, browptr, bcol, bval} (int64 / int32 / float32|float64)."""
        lib = _lib.load()
        self._lib = lib
        self._h = C.c_void_p(0)
        beam_ptr = np.ascontiguousarray(beam_ptr, dtype=np.int32)
        self.nb = len(beam_ptr) - 1
        self.beam_ptr = beam_ptr
        self.B = int(beam_ptr[-1])
        opt = Options()
        lib.vmat_default_options(C.byref(opt))
        opt.store_dtype = _lib.DTYPE_CODES[store]
        opt.acc_dtype = _lib.DTYPE_CODES[acc]
        opt.tile_voxels, opt.k1_threads, opt.k2_threads, opt.k2_tasks = tile_voxels, k1_threads, k2_threads, k2_tasks
        opt.use_graph = 1 if use_graph else 0
        opt.k1_stages, opt.k2_stages = k1_stages, k2_stages          # ring depths (0 = default)
        opt.chunk_cap, opt.k2_chunk_cap = chunk_cap, k2_chunk_cap    # chunk caps of the dose / transposed layouts
        opt.no_pdl = 0 if pdl else 1
        opt.host_path = host_path
        opt.claim_lead = claim_lead            # dose kernel: stages before a super-slice's end at which the next is claimed
        opt.index_bytes = index_bytes          # 4 int32 column indices, 2 uint16 whenever they fit, 1 byte deltas; 0 = byte deltas for fp32/fp32 plans of 5M nonzeros and more, else as 2
        opt.eval_kernels = fused               # 0 / 1 = the three kernels K1/K2/K3 (default), 2 = the single fused evaluation kernel
        opt.k2_pool_pct = k2_pool_pct          # transposed kernel: percent of the work claimed dynamically at the end (0 = default, -1 = none)
        opt.consume_batch = consume_batch      # ring stages a consumer warp takes per pass (0 = default: 2)
        self.store, self.acc = store, acc
        thr = np.ascontiguousarray(thr, dtype=np.float64)
        over = np.ascontiguousarray(over, dtype=np.float64)
        under = np.ascontiguousarray(under, dtype=np.float64)
        if isinstance(D, dict):
            keep = [D[k] for k in ("vrowptr", "vcol", "vval", "browptr", "bcol", "bval")]
            self.V = int(keep[0].numel()) - 1
            src = _lib.VMAT_F64 if str(keep[2].dtype).endswith("float64") else _lib.VMAT_F32
            loc = _lib.VMAT_DEVICE
        else:
            import scipy.sparse as sp
            Dr = sp.csr_matrix(D)
            Dr.sort_indices()
            Dc = Dt if Dt is not None else Dr.tocsc()
            Dc.sort_indices()
            self.V = Dr.shape[0]
            if Dr.shape[1] != self.B:
                raise VmatError(f"D has {Dr.shape[1]} columns but beam_ptr ends at {self.B}")
            vt = np.float64 if store == "f64" else np.float32
            keep = [np.ascontiguousarray(Dr.indptr, dtype=np.int64), np.ascontiguousarray(Dr.indices, dtype=np.int32),
                    np.ascontiguousarray(Dr.data, dtype=vt),
                    np.ascontiguousarray(Dc.indptr, dtype=np.int64), np.ascontiguousarray(Dc.indices, dtype=np.int32),
                    np.ascontiguousarray(Dc.data, dtype=vt)]
            src = _lib.VMAT_F64 if vt is np.float64 else _lib.VMAT_F32
            loc = _lib.VMAT_HOST
        if len(thr) != self.V:
            raise VmatError("penalty tables must have one entry per voxel")
        self.device = device
        check(lib.vmat_plan_create(C.byref(self._h), device, self.V, self.B, self.nb, _ptr(beam_ptr),
                                   *[_ptr(k) for k in keep], src, loc, _ptr(thr), _ptr(over), _ptr(under),
                                   C.byref(opt)), "vmat_plan_create")
        del keep
        self._y = np.zeros(self.nb)
        self._g = np.zeros(self.nb)
        self._f = C.c_double()
        # argument objects of the per-evaluation call, built once (vmat_eval is called ~10^4 times/s)
        self._eval_args = (self._h, _ptr(self._y), C.byref(self._f), _ptr(self._g))

    # -- lifecycle ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.vmat_plan_destroy(self._h)
            self._h.value = None            # in place: the prebuilt argument tuples see the null handle too

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        nnz, lay, alg, nl, iw = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int32()
        check(self._lib.vmat_plan_info(self._h, C.byref(nnz), C.byref(lay), C.byref(alg), C.byref(nl), C.byref(iw)),
              "vmat_plan_info")
        return dict(nnz=nnz.value, layout_bytes_per_eval=lay.value, algorithmic_bytes_per_eval=alg.value,
                    launches_per_eval=nl.value, index_bytes=iw.value, V=self.V, B=self.B, nb=self.nb, store=self.store, acc=self.acc)

    # -- aperture weights -----------------------------------------------------------
    def set_aperture(self, beam: int, w_dose, w_grad):
        """w_dose / w_grad: f64[B_beam] (strengths scattered over the open map / diagmaker)."""
        if not 0 <= beam < self.nb:
            raise VmatError(f"no control point {beam} (numbeams = {self.nb})")
        n = int(self.beam_ptr[beam + 1] - self.beam_ptr[beam])
        wd = None if w_dose is None else np.ascontiguousarray(w_dose, dtype=np.float64)
        wg = None if w_grad is None else np.ascontiguousarray(w_grad, dtype=np.float64)
        for w in (wd, wg):
            if w is not None and len(w) != n:
                raise VmatError(f"beam {beam} has {n} beamlets, got a weight vector of {len(w)}")
        check(self._lib.vmat_set_aperture(self._h, beam, _ptr(wd), _ptr(wg)), "vmat_set_aperture")

    def set_weights(self, w_dose, w_grad):
        """All control points at once from global f64[B] vectors."""
        for i in range(self.nb):
            lo, hi = int(self.beam_ptr[i]), int(self.beam_ptr[i + 1])
            self.set_aperture(i, w_dose[lo:hi], w_grad[lo:hi])

    # -- evaluation -----------------------------------------------------------------
    def _check(self, rc, what):
        """check() for the calls that run the evaluation kernels: a failure caused by the index guard carries its record."""
        if rc == 0:
            return
        try:
            check(rc, what)
        except VmatError as e:
            guard = None
            try:
                guard = self.guard_status()
            except VmatError:
                pass
            if guard:
                raise VmatError(f"{e}; index guard: {guard}") from None
            raise

    def eval(self, y):
        """(f, g[nb]) for intensities y[nb]: host buffers in, host buffers out."""
        self._y[:] = y
        rc = self._lib.vmat_eval(*self._eval_args)
        if rc:
            self._check(rc, "vmat_eval")
        return self._f.value, self._g.copy()

    def set_intensities(self, y):
        self._y[:] = y
        check(self._lib.vmat_set_intensities(self._h, _ptr(self._y)), "vmat_set_intensities")

    def eval_async(self, stream: Optional[int] = None, phase: int = 0):
        self._check(self._lib.vmat_eval_async(self._h, C.c_void_p(stream or 0), phase), "vmat_eval_async")

    def fetch_result(self):
        f = C.c_double()
        self._check(self._lib.vmat_fetch_result(self._h, C.byref(f), _ptr(self._g)), "vmat_fetch_result")
        return f.value, self._g.copy()

    def sync(self):
        self._check(self._lib.vmat_sync(self._h), "vmat_sync")

    def guard_status(self):
        """The index guard's record (vmat_guard_status): None while it has not tripped, else a dict.  Readable
        after a failed evaluation too (the record lives in pinned host memory)."""
        w = np.zeros(8, dtype=np.uint64)
        check(self._lib.vmat_guard_status(self._h, _ptr(w), len(w)), "vmat_guard_status")
        if not w[0]:
            return None
        hi = lambda k: int(w[k]) >> 32
        lo = lambda k: int(w[k]) & 0xffffffff
        return dict(kernel={1: "k1_dose_penalty", 2: "k2_beamlet_grad"}.get(hi(1), hi(1)), cta=lo(1), thread=hi(2), ring_slot=lo(2),
                    super_slice=hi(3), data_stage=lo(3), index_on_entry=hi(4), index_after=lo(4), limit=lo(5))

    def set_timing(self, enable: bool = True):
        check(self._lib.vmat_set_timing(self._h, 1 if enable else 0), "vmat_set_timing")

    def get_timing(self) -> dict:
        """Mean per-kernel milliseconds of the evaluations launched since set_timing(True)."""
        ms = np.zeros(4)
        n = C.c_int32()
        check(self._lib.vmat_get_timing(self._h, _ptr(ms), C.byref(n)), "vmat_get_timing")
        return dict(k1_ms=ms[0], k2_ms=ms[1], k3_ms=ms[2], eval_ms=ms[3], n=n.value)

    def trace_eval(self):
        """One evaluation with per-CTA time stamps: (k1[ctas, 5], k2[ctas, 5]) in ns relative to the
        earliest stamp; columns = entry, inputs ready, first stage landed, consumers done, producer done."""
        buf = np.zeros(8 * 4096 * 40, dtype=np.uint64)
        n1, n2 = C.c_int32(), C.c_int32()
        check(self._lib.vmat_trace_eval(self._h, _ptr(buf), buf.size, C.byref(n1), C.byref(n2)), "vmat_trace_eval")
        self.trace_items = None
        if n2.value == 0:       # fused kernel: per-item stamps of every CTA follow the per-CTA block
            it = buf[8 * n1.value: 8 * n1.value + n1.value * 32 * 8].reshape(n1.value, 32, 8).astype(np.int64)
            self.trace_items = it
        full = buf[:8 * (n1.value + n2.value)].reshape(-1, 8).astype(np.int64)
        t0 = full[full > 0].min()
        full = np.where(full > 0, full - t0, -1)
        self.trace_raw = full            # all 8 slots (the fused kernel uses 7: see tools/fused_probe.py)
        t = full[:, :5]
        self.trace_k3 = None
        if n2.value > 0:                 # three-kernel path: the reduction kernel's CTAs follow (one per control point)
            k3 = buf[8 * (n1.value + n2.value): 8 * (n1.value + n2.value + self.nb)].reshape(-1, 8).astype(np.int64)
            self.trace_k3 = np.where(k3 > 0, k3 - t0, -1)[:, :4]
        return t[:n1.value], t[n1.value:]

    def reduce_buffer(self):
        """torch view (float64, B+1: gb then f) of the buffer to all-reduce between phases 1 and 2."""
        import torch
        p, n = C.c_void_p(), C.c_int64()
        check(self._lib.vmat_reduce_buffer(self._h, C.byref(p), C.byref(n)), "vmat_reduce_buffer")
        return torch.as_tensor(_DeviceArray(p.value, n.value), device=f"cuda:{self.device}")

    def comm_handle(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's exchange buffer."""
        buf = (C.c_ubyte * 64)()
        check(self._lib.vmat_comm_handle(self._h, buf), "vmat_comm_handle")
        return bytes(buf)

    def comm_attach(self, rank: int, nranks: int, handles: bytes):
        """Map every rank's exchange buffer; from now on evaluations all-reduce over peer memory
        inside the reduction kernel.  Every rank must attach before any rank evaluates."""
        if len(handles) != 64 * nranks:
            raise VmatError("need one 64-byte IPC handle per rank")
        buf = (C.c_ubyte * len(handles)).from_buffer_copy(handles)
        check(self._lib.vmat_comm_attach(self._h, rank, nranks, buf), "vmat_comm_attach")
        self.fused_exchange = nranks > 1

    def comm_attach_local(self, rank: int, plans):
        """Same as comm_attach for plans of this process (`plans[rank] is self`): same device (tests) or
        devices with peer access."""
        arr = (C.c_void_p * len(plans))(*[p._h.value for p in plans])
        check(self._lib.vmat_comm_attach_local(self._h, rank, len(plans), arr), "vmat_comm_attach_local")
        self.fused_exchange = len(plans) > 1

    def comm_set_timeout(self, seconds: float):
        check(self._lib.vmat_comm_set_timeout(self._h, float(seconds)), "vmat_comm_set_timeout")

    def _device_view(self, fn):
        import torch
        p, n = C.c_void_p(), C.c_int64()
        check(fn(self._h, C.byref(p), C.byref(n)), "device buffer")
        return torch.as_tensor(_DeviceArray(p.value, n.value), device=f"cuda:{self.device}")

    def intensities_buffer(self):
        """torch view (float64[nb]) of the intensities the next eval_async reads."""
        return self._device_view(self._lib.vmat_intensities_buffer)

    def result_buffer(self):
        """torch view (float64[nb+1]: f then g) of what the last eval_async wrote."""
        return self._device_view(self._lib.vmat_result_buffer)

    def dose(self) -> np.ndarray:
        out = np.empty(self.V)
        check(self._lib.vmat_get_dose(self._h, _ptr(out)), "vmat_get_dose")
        return out

    def voxelgrad(self) -> np.ndarray:
        out = np.empty(self.V)
        check(self._lib.vmat_get_voxelgrad(self._h, _ptr(out)), "vmat_get_voxelgrad")
        return out

    def beamlet_grad(self) -> np.ndarray:
        out = np.empty(self.B)
        check(self._lib.vmat_get_beamlet_grad(self._h, _ptr(out)), "vmat_get_beamlet_grad")
        return out

    # -- restricted master problem: dense column cache ---------------------------------
    def rmp_begin(self, beams):
        b = np.ascontiguousarray(beams, dtype=np.int32)
        check(self._lib.vmat_rmp_begin(self._h, len(b), _ptr(b)), "vmat_rmp_begin")

    def rmp_eval(self, y):
        self._y[:] = y
        f = C.c_double()
        check(self._lib.vmat_rmp_eval(self._h, _ptr(self._y), C.byref(f), _ptr(self._g)), "vmat_rmp_eval")
        return f.value, self._g.copy()

    def rmp_end(self):
        check(self._lib.vmat_rmp_end(self._h), "vmat_rmp_end")

    # -- dose-volume histograms ------------------------------------------------------
    def dose_max(self) -> float:
        m = C.c_double()
        check(self._lib.vmat_dose_max(self._h, C.byref(m)), "vmat_dose_max")
        return m.value

    def dvh_histogram(self, full_mask, nstructs: int, edges) -> np.ndarray:
        """int64[nstructs, len(edges)-1]: voxels of structure s (bit s of full_mask) per dose bin."""
        fm = np.ascontiguousarray(full_mask, dtype=np.uint64)
        ed = np.ascontiguousarray(edges, dtype=np.float64)
        if len(fm) != self.V:
            raise VmatError("full_mask must have one entry per voxel")
        hist = np.zeros((nstructs, len(ed) - 1), dtype=np.int64)
        check(self._lib.vmat_dvh_histogram(self._h, _ptr(fm), nstructs, _ptr(ed), len(ed), _ptr(hist)), "vmat_dvh_histogram")
        return hist

    # -- pricing --------------------------------------------------------------------
    def price(self, cands: Sequence[int], params: PriceParams, rows):
        """Batched PPsubroutine: `rows` is a ctypes array of PriceRow, len(cands)*M."""
        n = len(cands)
        cand = np.ascontiguousarray(cands, dtype=np.int32)
        p = np.empty(n)
        l = np.empty((n, params.M), dtype=np.int32)
        r = np.empty((n, params.M), dtype=np.int32)
        check(self._lib.vmat_price(self._h, n, _ptr(cand), C.byref(params), rows, _ptr(p), _ptr(l), _ptr(r)),
              "vmat_price")
        return p, l, r


def _price_ex(self, cands, params, rows):
    """Batched PPsubroutine of any generation (params.variant): -> (p, [l...], [r...]) with one leaf list per
    candidate, as long as the reference returns it."""
    n = len(cands)
    cand = np.ascontiguousarray(cands, dtype=np.int32)
    p = np.empty(n)
    stride = params.M + 1
    l = np.empty((n, stride), dtype=np.int32)
    r = np.empty((n, stride), dtype=np.int32)
    ln = np.empty(n, dtype=np.int32)
    check(self._lib.vmat_price_ex(self._h, n, _ptr(cand), C.byref(params), rows, _ptr(p), _ptr(l), _ptr(r), _ptr(ln)),
          "vmat_price_ex")
    return p, [l[k, :ln[k]].tolist() for k in range(n)], [r[k, :ln[k]].tolist() for k in range(n)]


DosePlan.price_ex = _price_ex


def kmedoid_costs(points, curmin, device: int = 0, candidates=None) -> np.ndarray:
    """cost[c] = sum_j min(curmin[j], dist(c, j))  (kmedoids.py:32-60) on the GPU.  With
    `candidates` (n_cand x dim) the sum runs over `points` (e.g. one rank's shard) only."""
    P = np.ascontiguousarray(points, dtype=np.float64)
    if P.ndim == 1:
        P = P.reshape(-1, 1)
    cm = np.ascontiguousarray(curmin, dtype=np.float64)
    lib = _lib.load()
    if candidates is None:
        out = np.empty(P.shape[0])
        check(lib.vmat_kmedoid_costs(device, P.shape[0], P.shape[1], _ptr(P), _ptr(cm), _ptr(out)), "vmat_kmedoid_costs")
        return out
    Cc = np.ascontiguousarray(candidates, dtype=np.float64).reshape(-1, P.shape[1])
    out = np.empty(Cc.shape[0])
    check(lib.vmat_kmedoid_costs_shard(device, Cc.shape[0], _ptr(Cc), P.shape[0], P.shape[1], _ptr(P), _ptr(cm),
                                       _ptr(out)), "vmat_kmedoid_costs_shard")
    return out


class KMedoidSession:
    """Points, candidates and the running nearest-medoid distance resident on the device between insertion steps
    (vmat_kmedoid_open / _insert / _costs_resident): a step costs one distance pass, not an upload of the point set."""

    def __init__(self, points, curmin, candidates, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p(0)
        self.device = device
        P = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, np.shape(candidates)[1])
        Cc = np.ascontiguousarray(candidates, dtype=np.float64)
        cm = np.ascontiguousarray(curmin, dtype=np.float64)
        self.n_cand, self.dim = Cc.shape
        check(self._lib.vmat_kmedoid_open(C.byref(self._h), device, P.shape[0], self.dim, _ptr(P), _ptr(cm),
                                          self.n_cand, _ptr(Cc)), "vmat_kmedoid_open")

    def insert(self, medoid):
        m = np.ascontiguousarray(medoid, dtype=np.float64)
        check(self._lib.vmat_kmedoid_insert(self._h, _ptr(m)), "vmat_kmedoid_insert")

    def costs(self) -> np.ndarray:
        out = np.empty(self.n_cand)
        check(self._lib.vmat_kmedoid_costs_resident(self._h, _ptr(out), None), "vmat_kmedoid_costs_resident")
        return out

    def costs_device(self):
        """torch view (float64[n_cand]) of the costs, left on the device (for an all-reduce)."""
        import torch
        p = C.c_void_p()
        check(self._lib.vmat_kmedoid_costs_resident(self._h, None, C.byref(p)), "vmat_kmedoid_costs_resident")
        return torch.as_tensor(_DeviceArray(p.value, self.n_cand), device=f"cuda:{self.device}")

    def close(self):
        if self._h.value:
            self._lib.vmat_kmedoid_close(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
