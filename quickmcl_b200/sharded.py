"""Particle sharding across GPUs: one process per GPU, each owning a
contiguous block of the global particle index range, `torch.distributed` for the
exchanges.

The sharded algorithm itself lives in the CUDA library (quickmcl_b200/csrc:
filter.cu, resample.cu, cluster.cu, shard.cu); it asks the host for three
collectives through the `qmcl_comm` table of include/qmcl.h.  `TorchComm` fills
that table with torch.distributed calls (NCCL over NVLink on GPUs; gloo on CPU
buffers for the host-logic tests), and `ShardedParticleFilter` is the
IParticleFilter mirror whose handle has the table installed.  The reference has
no multi-process code at all; the cycle order is still laser_handler.cpp:243-276.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _abi
from ._abi import check, ptr
from .api import PARTICLE_DTYPE, Particle, ParticleFilter

_DT = {0: torch.int32, 1: torch.int64, 2: torch.float64}
_OP = {0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                           C.c_void_p)
ALLTOALLV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                           C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64),
                           C.POINTER(C.c_uint64), C.c_void_p)


class QmclComm(C.Structure):
    """qmcl_comm (include/qmcl.h)."""
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("ctx", C.c_void_p),
                ("allgather", ALLGATHER_FN), ("allreduce", ALLREDUCE_FN),
                ("alltoallv", ALLTOALLV_FN)]


class _DeviceBytes:
    """Zero-copy view of raw device memory for torch (CUDA array interface v2)."""

    def __init__(self, address, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(address), False), "version": 2}


class TorchComm:
    """The three collectives of qmcl_comm on top of a torch.distributed group.

    device="cuda": buffers are device pointers, calls are enqueued behind the
    given CUDA stream (NCCL).  device="cpu": buffers are host pointers (gloo);
    used by the CPU tests of the multi-rank host logic.
    """

    def __init__(self, group=None, device="cuda"):
        self.group = group if group is not None else dist.group.WORLD
        self.device = device
        self.rank = dist.get_rank(self.group)
        self.size = dist.get_world_size(self.group)
        self.calls = {"allgather": 0, "allreduce": 0, "alltoallv": 0}
        self.error = None
        # the ctypes thunks must outlive the filter that uses them
        self._ag = ALLGATHER_FN(self._allgather)
        self._ar = ALLREDUCE_FN(self._allreduce)
        self._a2a = ALLTOALLV_FN(self._alltoallv)
        self.struct = QmclComm(self.rank, self.size, None, self._ag, self._ar, self._a2a)

    # -- raw memory -> tensor ------------------------------------------------
    def _bytes(self, address, nbytes):
        if nbytes == 0:
            return torch.empty(0, dtype=torch.uint8, device=self.device)
        if self.device == "cpu":
            arr = np.ctypeslib.as_array((C.c_uint8 * int(nbytes)).from_address(int(address)))
            return torch.from_numpy(arr)
        return torch.as_tensor(_DeviceBytes(address, nbytes), device="cuda")

    def _on_stream(self, stream):
        if self.device == "cpu":
            import contextlib
            return contextlib.nullcontext()
        if not stream:
            return torch.cuda.stream(torch.cuda.default_stream())
        return torch.cuda.stream(torch.cuda.ExternalStream(int(stream)))

    def _guard(self, fn, *a):
        try:
            fn(*a)
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the C status
            self.error = e
            return 1

    # -- the callbacks -------------------------------------------------------
    def _allgather(self, ctx, buf, bytes_per_rank, stream):
        def run():
            self.calls["allgather"] += 1
            whole = self._bytes(buf, bytes_per_rank * self.size)
            mine = whole[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank].clone()
            with self._on_stream(stream):
                if self.device == "cpu":
                    parts = [torch.empty_like(mine) for _ in range(self.size)]
                    dist.all_gather(parts, mine, group=self.group)
                    whole.copy_(torch.cat(parts))
                else:
                    dist.all_gather_into_tensor(whole, mine, group=self.group)
        return self._guard(run)

    def _allreduce(self, ctx, buf, count, dtype, op, stream):
        def run():
            self.calls["allreduce"] += 1
            dt = _DT[dtype]
            t = self._bytes(buf, count * dt.itemsize).view(dt)
            with self._on_stream(stream):
                dist.all_reduce(t, op=_OP[op], group=self.group)
        return self._guard(run)

    def _alltoallv(self, ctx, send, send_bytes, send_off, recv, recv_bytes, recv_off, stream):
        def run():
            self.calls["alltoallv"] += 1
            sb = [int(send_bytes[i]) for i in range(self.size)]
            so = [int(send_off[i]) for i in range(self.size)]
            rb = [int(recv_bytes[i]) for i in range(self.size)]
            ro = [int(recv_off[i]) for i in range(self.size)]

            def packed(sizes, offs):
                # all_to_all_single wants the per-peer pieces back to back in rank order
                live = [(o, s) for o, s in zip(offs, sizes) if s]
                start = live[0][0] if live else 0
                at = start
                for o, s in live:
                    if o != at:
                        raise ValueError("alltoallv pieces are not contiguous in rank order")
                    at += s
                return start, at - start
            s0, sn = packed(sb, so)
            r0, rn = packed(rb, ro)
            src = self._bytes(send + s0 if send else 0, sn)
            dst = self._bytes(recv + r0 if recv else 0, rn)
            with self._on_stream(stream):
                dist.all_to_all_single(dst, src, output_split_sizes=rb, input_split_sizes=sb,
                                       group=self.group)
        return self._guard(run)


def shard_slice(n, rank, size):
    """Block [first, first+count) of `rank` (qmcl_shard_slice)."""
    first, count = C.c_uint64(0), C.c_uint64(0)
    _abi.lib().qmcl_shard_slice(n, rank, size, C.byref(first), C.byref(count))
    return first.value, count.value


def lv_owner_range(M, r, minv, c_lo, c_hi, is_first, is_last):
    """Output slots [m_lo, m_hi) owned by a shard whose CDF spans (c_lo, c_hi]."""
    lo, hi = C.c_uint64(0), C.c_uint64(0)
    _abi.lib().qmcl_lv_owner_range(M, r, minv, c_lo, c_hi, int(is_first), int(is_last),
                                   C.byref(lo), C.byref(hi))
    return lo.value, hi.value


class ShardedParticleFilter(ParticleFilter):
    """IParticleFilter over a particle set sharded across the ranks of `group`.

    Every rank constructs one with the same map, seed and parameters and makes
    the same calls in the same order (SPMD).  `num_particles()` is the global
    count; `get_particles()` returns this rank's block and `gather_particles()`
    the whole set in global order.
    """

    def __init__(self, map_, seed=0, group=None, transport="nccl"):
        """transport="nccl": the library's built-in NCCL table (qmcl_nccl_comm_create; the
        process group only carries the 128-byte unique id once).  transport="torch": every
        collective goes through torch.distributed (TorchComm)."""
        super().__init__(map_, seed)
        self.group = group if group is not None else dist.group.WORLD
        self._rank = dist.get_rank(self.group)
        self._size = dist.get_world_size(self.group)
        self.comm = None
        self._nccl = None
        if transport == "torch":
            self.comm = TorchComm(self.group, "cuda")
            check(self._L.qmcl_filter_set_comm(self._h, C.byref(self.comm.struct)),
                  "qmcl_filter_set_comm")
        elif transport == "nccl":
            uid = (C.c_uint8 * 128)()
            if self._rank == 0:
                check(self._L.qmcl_nccl_unique_id(uid), "qmcl_nccl_unique_id")
            box = [bytes(uid)]
            dist.broadcast_object_list(box, src=dist.get_global_rank(self.group, 0),
                                       group=self.group)
            uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
            self._nccl = QmclComm()
            check(self._L.qmcl_nccl_comm_create(uid, self._rank, self._size,
                                                torch.cuda.current_device(),
                                                C.byref(self._nccl)), "qmcl_nccl_comm_create")
            check(self._L.qmcl_filter_set_comm(self._h, C.byref(self._nccl)),
                  "qmcl_filter_set_comm")
        else:
            raise ValueError(transport)

    def close(self):
        if getattr(self, "_nccl", None) is not None and getattr(self, "_h", None):
            self._L.qmcl_filter_set_comm(self._h, None)
            self._L.qmcl_nccl_comm_destroy(C.byref(self._nccl))
            self._nccl = None
        super().close()

    @property
    def rank(self):
        return self._rank

    @property
    def world_size(self):
        return self._size

    def set_rebalance_threshold(self, relative_imbalance):
        """Skip the rebalancing all-to-all while every rank is within this fraction of its share."""
        check(self._L.qmcl_filter_set_rebalance_threshold(self._h, float(relative_imbalance)),
              "qmcl_filter_set_rebalance_threshold")

    def set_rebalance(self, enabled):
        check(self._L.qmcl_filter_set_rebalance(self._h, int(enabled)),
              "qmcl_filter_set_rebalance")

    def shard_range(self):
        """(first global index, local count, global count)."""
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        check(self._L.qmcl_filter_shard_range(self._h, C.byref(a), C.byref(b), C.byref(c)),
              "qmcl_filter_shard_range")
        return a.value, b.value, c.value

    def shard_stats(self):
        """Exchange counters since creation: collectives enqueued, carry-chain rounds that
        needed an exchange of their own, rebalancing all-to-alls skipped."""
        st = (C.c_uint64 * 3)()
        check(self._L.qmcl_filter_shard_stats(self._h, st), "qmcl_filter_shard_stats")
        return {"collectives": int(st[0]), "chain_rounds": int(st[1]), "rebalances_skipped": int(st[2])}

    def set_particles(self, particles):
        """`particles` is the GLOBAL set (same array on every rank); each rank keeps its block."""
        particles = np.ascontiguousarray(particles)
        assert particles.dtype == PARTICLE_DTYPE
        n = len(particles)
        first, count = shard_slice(n, self.rank, self.world_size)
        mine = np.ascontiguousarray(particles[first:first + count])
        check(self._L.qmcl_filter_set_particles_shard(self._h, ptr(mine, Particle), count, first,
                                                      n), "qmcl_filter_set_particles_shard")

    def get_particles(self):
        _, count, _ = self.shard_range()
        out = np.zeros(count, dtype=PARTICLE_DTYPE)
        if count:
            check(self._L.qmcl_filter_copy_particles(self._h, ptr(out, Particle), count),
                  "qmcl_filter_copy_particles")
        return out

    def gather_particles(self):
        """The whole set in global index order, on every rank (test/diagnostic use)."""
        mine = self.get_particles()
        first, count, total = self.shard_range()
        meta = [None] * self.world_size
        dist.all_gather_object(meta, (first, mine), group=self.group)
        meta.sort(key=lambda m: m[0])
        out = np.concatenate([m[1] for m in meta]) if meta else mine
        assert len(out) == total, (len(out), total)
        return out

    def _raise_comm(self):
        if self.comm is not None and self.comm.error is not None:
            e, self.comm.error = self.comm.error, None
            raise e


class LocalComm:
    """qmcl_comm for shards that are THREADS of one process on one GPU.

    Testing aid: lets the sharded code path of the library run (and be checked
    against the oracle) on a single-GPU box — each thread drives one shard
    handle and the collectives are device-to-device copies between the shards'
    buffers, ordered with host-side barriers (no kernel ever waits on another).
    Production runs use TorchComm (one process per GPU, NCCL).
    """

    class Shared:
        def __init__(self, size):
            import threading
            self.size = size
            self.barrier = threading.Barrier(size)
            self.slots = [None] * size

    def __init__(self, shared, rank):
        self.shared = shared
        self.rank = rank
        self.size = shared.size
        self.calls = {"allgather": 0, "allreduce": 0, "alltoallv": 0}
        self.error = None
        self._ag = ALLGATHER_FN(self._allgather)
        self._ar = ALLREDUCE_FN(self._allreduce)
        self._a2a = ALLTOALLV_FN(self._alltoallv)
        self.struct = QmclComm(self.rank, self.size, None, self._ag, self._ar, self._a2a)

    @staticmethod
    def _bytes(address, nbytes):
        if nbytes == 0:
            return torch.empty(0, dtype=torch.uint8, device="cuda")
        return torch.as_tensor(_DeviceBytes(address, nbytes), device="cuda")

    def _exchange(self, mine):
        """Publish `mine`, wait for everybody, return all ranks' objects."""
        sh = self.shared
        torch.cuda.synchronize()
        sh.slots[self.rank] = mine
        sh.barrier.wait()
        everyone = list(sh.slots)
        return everyone

    def _done(self):
        torch.cuda.synchronize()
        self.shared.barrier.wait()

    def _guard(self, fn):
        try:
            fn()
            return 0
        except Exception as e:  # noqa: BLE001
            self.error = e
            try:
                self.shared.barrier.abort()
            except Exception:  # noqa: BLE001
                pass
            return 1

    def _allgather(self, ctx, buf, bytes_per_rank, stream):
        def run():
            self.calls["allgather"] += 1
            whole = self._bytes(buf, bytes_per_rank * self.size)
            everyone = self._exchange(whole)
            for r, other in enumerate(everyone):
                if r != self.rank:
                    sl = slice(r * bytes_per_rank, (r + 1) * bytes_per_rank)
                    whole[sl].copy_(other[sl])
            self._done()
        return self._guard(run)

    def _allreduce(self, ctx, buf, count, dtype, op, stream):
        def run():
            self.calls["allreduce"] += 1
            dt = _DT[dtype]
            t = self._bytes(buf, count * dt.itemsize).view(dt)
            everyone = self._exchange(t)
            stack = torch.stack(everyone)
            res = {0: stack.sum(0), 1: stack.min(0).values, 2: stack.max(0).values}[op]
            self._done()          # everybody has read before anybody overwrites
            t.copy_(res)
            self._done()
        return self._guard(run)

    def _alltoallv(self, ctx, send, send_bytes, send_off, recv, recv_bytes, recv_off, stream):
        def run():
            self.calls["alltoallv"] += 1
            sb = [int(send_bytes[i]) for i in range(self.size)]
            so = [int(send_off[i]) for i in range(self.size)]
            total = max([o + s for o, s in zip(so, sb)] + [0])
            src = self._bytes(send, total)
            everyone = self._exchange((src, sb, so))
            for r, (osrc, osb, oso) in enumerate(everyone):
                nb = int(recv_bytes[r])
                assert nb == osb[self.rank], (nb, osb[self.rank])
                if nb:
                    dst = self._bytes(recv + int(recv_off[r]), nb)
                    dst.copy_(osrc[oso[self.rank]:oso[self.rank] + nb])
            self._done()
        return self._guard(run)


def run_local_shards(n_shards, make_map, body, seed=0):
    """Run `body(filter, rank)` on `n_shards` shard handles, one thread each, on the
    current GPU; returns the list of results in rank order.  `make_map()` builds
    one map per shard (maps are not shared between threads)."""
    import threading
    shared = LocalComm.Shared(n_shards)
    results = [None] * n_shards
    errors = []

    def work(rank):
        try:
            gm = make_map()
            f = ParticleFilter(gm, seed)
            comm = LocalComm(shared, rank)
            f._local_comm = comm
            check(f._L.qmcl_filter_set_comm(f._h, C.byref(comm.struct)), "qmcl_filter_set_comm")
            results[rank] = body(f, rank)
        except Exception as e:  # noqa: BLE001
            errors.append((rank, e))
            try:
                shared.barrier.abort()
            except Exception:  # noqa: BLE001
                pass

    threads = [threading.Thread(target=work, args=(r,)) for r in range(n_shards)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        # prefer the root cause over the BrokenBarrierError it triggered elsewhere
        import threading as _t
        errors.sort(key=lambda e: isinstance(e[1], _t.BrokenBarrierError))
        raise errors[0][1]
    return results
