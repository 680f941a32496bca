"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), contiguous slices of the
flattened source-position / polynomial axis, NO collective on the data path, one final gather.

Every polynomial (cpu_ops.cc:45-72) and every source position (lightcurve.py:245-254) is an
independent unit, so the only communication is the result gather."""
import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "sharded_apply", "balanced_order", "row_block", "sharded_rows", "PeerGather",
           "HostGather"]


def shard_bounds(n, world, rank):
    """contiguous ceil(n / world) split; the last ranks may get shorter (or empty) slices"""
    per = -(-n // world) if world > 0 else n
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def balanced_order(cost, world):
    """Permutation that deals the units round-robin in decreasing-cost order, so that gated workloads
    (a full contour integration costs ~10^3 hexadecapole points) spread evenly over contiguous
    shards.  Returns (perm, inverse): apply fn to x[perm] sharded, then out[inverse]."""
    cost = np.asarray(cost)
    order = np.argsort(-cost, kind="stable")
    n = len(order)
    per = -(-n // world)
    # position k of rank r's shard <- the (k*world + r)-th most expensive unit
    perm = np.full(per * world, -1, dtype=np.int64)
    for r in range(world):
        sel = order[r::world]
        perm[r * per:r * per + len(sel)] = sel
    perm = perm[perm >= 0]
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    return perm, inv


def sharded_apply(fn, x, group=None, gather=True):
    """Apply `fn` (rows in -> rows out, same leading length) to this rank's slice of `x` and, if
    `gather`, return the full result on every rank (one all_gather; padded to equal shard sizes).
    Works without an initialised process group (single process)."""
    if not (dist.is_available() and dist.is_initialized()):
        return fn(x)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = x.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    local = fn(x[lo:hi])
    if not gather:
        return local
    per = -(-n // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n]


def row_block(nrows, world, rank, align=1):
    """Row block [lo, hi) of rank `rank` for a map of `nrows` rows: contiguous, equal to within one
    `align`-row unit.  With align = 32 (the walk length of csrc/ps_walk.cuh) every rank's block starts on a
    walk boundary of the unsharded map, so a map computed with `walk=True` does not depend on the number of
    ranks as long as every block is large enough to be given full 32-row walks (the launcher shortens the
    walks of maps too small to fill the GPU); each column segment is then the same walk wherever it runs.
    The default (align = 1) balances the rows exactly instead."""
    units = -(-nrows // align)
    per = -(-units // world) if world > 0 else units
    lo = min(rank * per * align, nrows)
    return lo, min(lo + per * align, nrows)


def sharded_rows(fn, nrows, group=None, gather=True, align=1):
    """Row-sharded evaluation of a map (BASELINE config C5): `fn(lo, hi)` returns this rank's rows as a
    (hi - lo, ...) tensor; no data-path collective, one all_gather at the end if `gather`.  Example:

        mag = sharded_rows(lambda lo, hi: mag_point_source_map(x0, y0, dx, dy, nx, ny, rows=(lo, hi), s=s, q=q),
                           ny, align=32)

    Works without an initialised process group (single process)."""
    if not (dist.is_available() and dist.is_initialized()):
        return fn(0, nrows)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = row_block(nrows, world, rank, align)
    local = fn(lo, hi)
    if not gather:
        return local
    per = row_block(nrows, world, 0, align)[1]
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:nrows]


class _RawCuda:
    """a raw device pointer as a __cuda_array_interface__ object (-> torch.as_tensor without a copy)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerGather:
    """The final result gather WITHOUT a collective (include/caustics_b200.h, caustics_peer_*).

    Rank `dst` owns one device buffer of `nbytes`; every other rank maps it over NVLink (CUDA IPC) and
    passes `ptr(offset)` as the result pointer of a launcher, so the kernels' own stores land in the
    destination while they compute -- there is no gather phase to wait for, only `finish()`: stream
    synchronise + barrier, after which `tensor(dtype, shape)` on the destination holds everyone's slice.
    Setup (allocation, handle exchange, mapping) happens once, outside any timed region; the buffer is
    reused across calls.  Without an initialised process group it degrades to a plain local buffer."""

    def __init__(self, nbytes, dst=0, group=None):
        import ctypes
        from . import _lib
        self._L, self._lib = _lib.lib(), _lib
        self.nbytes, self.dst, self.group = int(nbytes), dst, group
        self.dist = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.dist else 0
        self.world = dist.get_world_size(group) if self.dist else 1
        self.owner = self.rank == dst or not self.dist
        self._base = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        if self.owner:
            _lib.check(self._L.caustics_peer_alloc(ctypes.byref(self._base), self.nbytes))
            if self.dist:
                _lib.check(self._L.caustics_peer_export(self._base, handle))
        if self.dist:
            box = [bytes(handle) if self.owner else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group is not None else dst,
                                       group=group)
            if not self.owner:
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(box[0])
                _lib.check(self._L.caustics_peer_open(buf, ctypes.byref(self._base)))
        self._open = True

    def ptr(self, offset_bytes=0):
        """device address of byte `offset_bytes` of the destination buffer, valid on THIS rank"""
        if not (0 <= offset_bytes <= self.nbytes):
            raise ValueError("offset outside the peer buffer")
        return self._base.value + int(offset_bytes)

    def tensor(self, dtype=torch.float64, shape=None):
        """the destination buffer as a torch tensor (no copy); meaningful on the owner after finish()"""
        t = torch.as_tensor(_RawCuda(self._base.value, self.nbytes), device="cuda").view(dtype)
        return t if shape is None else t[: int(np.prod(shape))].reshape(shape)

    def finish(self):
        """every rank's stores are complete and visible on the destination when this returns"""
        torch.cuda.current_stream().synchronize()
        if self.dist:
            dist.barrier(group=self.group)

    def close(self):
        if not self._open:
            return
        self._open = False
        if self.dist:
            dist.barrier(group=self.group)          # nobody is still writing
        if self.owner:
            self._L.caustics_peer_free(self._base)
        else:
            self._L.caustics_peer_close(self._base)

    def __del__(self):
        try:
            if self._open and not self.dist:
                self._L.caustics_peer_free(self._base)
        except Exception:
            pass


class HostGather:
    """The same for HOST results (the end-to-end path of a one-process-per-GPU driver): rank `dst` creates
    one POSIX shared-memory segment, every rank maps and page-locks it (cudaHostRegister), and the *_host
    entry points write their slices into it with their own D2H copies -- eight PCIe links in parallel
    instead of a device-side gather followed by one 800 MB copy over a single link."""

    def __init__(self, nbytes, dst=0, group=None):
        from multiprocessing import shared_memory
        self.nbytes, self.dst, self.group = int(nbytes), dst, group
        self.dist = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.dist else 0
        self.owner = self.rank == dst or not self.dist
        box = [None]
        if self.owner:
            self.shm = shared_memory.SharedMemory(create=True, size=self.nbytes)
            box[0] = self.shm.name
        if self.dist:
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, dst) if group is not None else dst,
                                       group=group)
            if not self.owner:
                self.shm = shared_memory.SharedMemory(name=box[0])
                try:        # Python < 3.13 registers attached segments for unlinking too: only the owner unlinks
                    from multiprocessing import resource_tracker
                    resource_tracker.unregister(self.shm._name, "shared_memory")
                except Exception:
                    pass
        self.array = np.ndarray((self.nbytes,), dtype=np.uint8, buffer=self.shm.buf)
        self.array[self.rank::4096] = 0            # fault the pages in before pinning
        self._registered = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.array.ctypes.data, self.nbytes, 0)
            self._registered = int(rc) == 0
        if self.dist:
            dist.barrier(group=group)

    def view(self, dtype=np.float64, shape=None):
        a = self.array.view(dtype)
        return a if shape is None else a[: int(np.prod(shape))].reshape(shape)

    def finish(self):
        if self.dist:
            dist.barrier(group=self.group)

    def close(self):
        if getattr(self, "shm", None) is None:
            return
        if self._registered:
            torch.cuda.cudart().cudaHostUnregister(self.array.ctypes.data)
        if self.dist:
            dist.barrier(group=self.group)
        self.array = None
        shm, self.shm = self.shm, None
        shm.close()
        if self.owner:
            shm.unlink()
