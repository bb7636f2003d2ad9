"""Device runtime: one libfrkb200 context per process/GPU.  PyTorch supplies device memory,
the stream and ``torch.distributed`` (NCCL) -- plumbing only; every computation on the hot
path is a libfrkb200 kernel reached through the C ABI with raw device pointers."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from ._lib import FRKError, call, lib, vp

_DEVICE = None


class Device:
    def __init__(self, index: Optional[int] = None):
        import torch
        if not torch.cuda.is_available():
            raise FRKError("frk_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if index is None:
            index = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device()))
        self.torch = torch
        self.index = index
        torch.cuda.set_device(index)
        self.tdev = torch.device("cuda", index)
        # a real (non-default) stream, made current for torch as well: libfrkb200 captures its factorisation launch sequences into
        # CUDA graphs, which the legacy default stream does not allow; torch allocations / NCCL calls stay ordered with the kernels
        self.stream = torch.cuda.Stream(device=index, priority=-1)
        torch.cuda.set_stream(self.stream)
        ctx = vp()
        call("frk_ctx_create", index, vp(self.stream.cuda_stream), C.byref(ctx))
        self.ctx = ctx
        self.has_comm = False

    # ---- in-library NCCL (frk_comm_*): no torch, no Python in the data path of the reductions -------------------
    def init_comm(self, rank: Optional[int] = None, world: Optional[int] = None, id_file: Optional[str] = None):
        """Give the context its own NCCL communicator.  The 128-byte unique id travels from rank 0 to the others either through the
        already initialised ``torch.distributed`` group (any backend -- rendezvous only) or, with ``id_file``, through a file
        (what an R / MPI host would do).  Models created afterwards reduce over the communicator instead of the Python hook."""
        import time as _time
        d = self.torch.distributed
        have_pg = d.is_available() and d.is_initialized()
        rank = (d.get_rank() if have_pg else 0) if rank is None else int(rank)
        world = (d.get_world_size() if have_pg else 1) if world is None else int(world)
        if world <= 1:
            return self
        uid = C.create_string_buffer(128)
        if id_file is not None:
            if rank == 0:
                call("frk_comm_unique_id", uid)
                with open(id_file + ".tmp", "wb") as fh:
                    fh.write(uid.raw)
                os.replace(id_file + ".tmp", id_file)
            else:
                t0 = _time.time()
                while not os.path.exists(id_file):
                    if _time.time() - t0 > 120:
                        raise FRKError(f"init_comm: no unique id at {id_file} after 120 s")
                    _time.sleep(0.01)
                with open(id_file, "rb") as fh:
                    uid.raw = fh.read(128)
        else:
            if not have_pg:
                raise FRKError("init_comm: needs an initialised torch.distributed group or id_file to share the NCCL unique id")
            box = [None]
            if rank == 0:
                call("frk_comm_unique_id", uid)
                box = [uid.raw]
            d.broadcast_object_list(box, src=0)
            uid.raw = box[0]
        call("frk_comm_init", self.ctx, rank, world, uid)
        self.has_comm = True
        self._comm_rank, self._comm_world = rank, world
        return self

    # ---- memory (torch-owned) ------------------------------------------------------
    def empty(self, n, dtype="f8"):
        t = self.torch
        dt = {"f8": t.float64, "i4": t.int32, "i8": t.int64, "u1": t.uint8}[dtype]
        return t.empty(int(n), dtype=dt, device=self.tdev)

    def zeros(self, n, dtype="f8"):
        x = self.empty(n, dtype)
        call("frk_memset0", self.ctx, self.ptr(x), x.numel() * x.element_size())
        return x

    def upload(self, a: np.ndarray, dtype=None):
        """Host array -> device (column-major flattening for 2-D inputs, as R stores matrices)."""
        a = np.asarray(a)
        if dtype is not None:
            a = a.astype(dtype, copy=False)
        flat = np.ascontiguousarray(a.T).ravel() if a.ndim == 2 else np.ascontiguousarray(a).ravel()
        code = {"float64": "f8", "int32": "i4", "int64": "i8"}[str(flat.dtype)]
        out = self.empty(flat.size, code)
        if flat.size:
            call("frk_h2d", self.ctx, self.ptr(out), vp(flat.ctypes.data), flat.nbytes)
            call("frk_sync", self.ctx)      # flat may be a temporary
        return out

    def host_buffer(self, n: int, dtype=np.float64) -> np.ndarray:
        """Pinned host array (the memory belongs to a torch tensor that the returned array keeps alive): D2H copies of the
        r x r slots and of the N-sized prediction vectors run at PCIe speed instead of through a pageable bounce buffer."""
        td = {np.float64: self.torch.float64, np.int32: self.torch.int32, np.int64: self.torch.int64}[dtype]
        return self.torch.empty(int(max(n, 1)), dtype=td, pin_memory=True).numpy()[:int(n)]

    def download(self, t, n: Optional[int] = None) -> np.ndarray:
        n = t.numel() if n is None else int(n)
        npdt = {8: {True: np.float64, False: np.int64}, 4: {False: np.int32}}[t.element_size()][t.is_floating_point()]
        out = self.host_buffer(n, npdt) if n * t.element_size() >= (1 << 20) else np.empty(n, dtype=npdt)
        if n:
            call("frk_d2h", self.ctx, vp(out.ctypes.data), self.ptr(t), out.nbytes)
        return out

    @staticmethod
    def ptr(t, offset_elems: int = 0):
        if t is None:
            return vp(None)
        return vp(t.data_ptr() + offset_elems * t.element_size())

    def set_deterministic(self, on: bool = True):
        """Run-to-run bit-reproducible results on one GPU (frk_use_deterministic): the Gram tiles are added in a fixed order
        instead of with fp64 atomics."""
        call("frk_use_deterministic", self.ctx, 1 if on else 0)

    def sync(self):
        call("frk_sync", self.ctx)

    def launches(self) -> int:
        return int(lib.frk_launch_count(self.ctx))

    # ---- distributed plumbing --------------------------------------------------------
    @property
    def world(self) -> int:
        if self.has_comm:
            return self._comm_world
        d = self.torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    @property
    def rank(self) -> int:
        if self.has_comm:
            return self._comm_rank
        d = self.torch.distributed
        return d.get_rank() if d.is_available() and d.is_initialized() else 0

    def allreduce_(self, t):
        """In-place fp64 sum over ranks (NCCL over NVLink); no-op for a single rank."""
        if self.world > 1:
            self.torch.distributed.all_reduce(t)
        return t

    def allreduce_hook(self):
        """ctypes callback handed to frk_model_create: libfrkb200 calls it for every cross-rank reduction.  Device
        buffers are wrapped zero-copy (``__cuda_array_interface__``) and reduced by NCCL on the ctx stream."""
        from ._lib import ALLREDUCE_FN
        torch, tdev = self.torch, self.tdev

        class _View:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}

        def hook(user, buf, n, on_device, op):
            try:
                d = torch.distributed
                rop = d.ReduceOp.SUM if op == 0 else d.ReduceOp.MAX
                nccl = d.get_backend() == "nccl"
                if on_device:
                    t = torch.as_tensor(_View(buf, n), device=tdev)
                    if nccl:
                        d.all_reduce(t, op=rop)
                    else:                       # gloo (several ranks sharing one GPU in the tests): through the host
                        hcpu = t.cpu()
                        d.all_reduce(hcpu, op=rop)
                        t.copy_(hcpu)
                else:
                    h = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(int(n),))
                    t = torch.as_tensor(h.copy(), device=tdev if nccl else "cpu")
                    d.all_reduce(t, op=rop)
                    h[:] = t.cpu().numpy()
                return 0
            except Exception as e:      # never let an exception cross the C ABI
                print(f"frk_b200 all-reduce hook failed: {e!r}", flush=True)
                return 1
        return ALLREDUCE_FN(hook)

    def allreduce_host(self, a: np.ndarray) -> np.ndarray:
        """Sum a small host vector over ranks (M-step moments, O(p^2) doubles)."""
        return allreduce_host(a, "sum", self.tdev)

    def allreduce_host_max(self, a: np.ndarray) -> np.ndarray:
        return allreduce_host(a, "max", self.tdev)


def allreduce_host(a: np.ndarray, op: str = "sum", device=None) -> np.ndarray:
    """All-reduce a small host vector over the ranks of the default process group (NCCL on the GPU
    box, gloo in the CPU tests); identity when torch.distributed is not initialised."""
    import torch
    d = torch.distributed
    a = np.ascontiguousarray(a, dtype=np.float64)
    if not (d.is_available() and d.is_initialized()) or d.get_world_size() == 1:
        return a
    on_gpu = d.get_backend() == "nccl"
    t = torch.as_tensor(a.copy(), device=device if on_gpu else "cpu")
    d.all_reduce(t, op=d.ReduceOp.SUM if op == "sum" else d.ReduceOp.MAX)
    return t.cpu().numpy()


def get_device(index: Optional[int] = None) -> Device:
    """Process-wide device runtime (created on first use)."""
    global _DEVICE
    if _DEVICE is None:
        _DEVICE = Device(index)
    return _DEVICE


def partition_rows_weighted(weights, world: int, rank: int):
    """Contiguous partition of len(weights) rows over `world` ranks with (nearly) equal weight per rank: the cut after rank k is
    the first row at which the cumulative weight reaches (k + 1) / world of the total.  Every rank computes the same cuts from
    the same weights.  Used when the work per BAU row is uneven (observations are not uniform over the BAU index)."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world <= 1 or n == 0:
        return 0, n
    c = np.cumsum(w)
    cuts = [0] + [int(np.searchsorted(c, c[-1] * (k + 1) / world, side="left")) + 1 for k in range(world - 1)] + [n]
    cuts = np.minimum.accumulate(np.asarray(cuts[::-1]))[::-1]                # (monotone even with long runs of zero weight)
    cuts = np.maximum.accumulate(np.clip(cuts, 0, n))
    return int(cuts[rank]), int(cuts[rank + 1])


def partition_rows(n: int, world: int, rank: int):
    """Contiguous block partition of n rows over `world` ranks: [lo, hi) for `rank`
    (BAUs block-partitioned by row range; each observation lives on the rank
    that owns its BAU)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi
