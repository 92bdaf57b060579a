"""world_size-2 gloo tests (CPU) of the host side of the sharded path:

* the `qmcl_comm` table as filled by quickmcl_b200.sharded.TorchComm, called
  through its C function pointers exactly as the CUDA library calls it;
* the pure host arithmetic the library exports for sharding
  (qmcl_shard_slice, qmcl_lv_owner_range);
* the sharded low-variance resampling protocol end to end — carry chain,
  owner-computes slot ranges, rebalancing all-to-all — emulated per rank with
  numpy standing in for the per-shard kernels, against the CPU oracle's
  resample_low_variance (particle_filter.cpp:214-239) on the whole set:
  the indices must be bit-exact.

No CUDA call is made here (there is no GPU in this container); the GPU side of
the same protocol is tests/test_sharded_gpu.py.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _u53(lo, hi):
    return float(((int(hi) << 32) | int(lo)) >> 11) * 2.0 ** -53


def _lv_offset(seed, step):
    """r of the shared RNG contract: Philox(key=seed, ctr=(0, 0, step, P_LV_OFFSET=4))."""
    from oracle import binding as ob
    ctr = (C.c_uint32 * 4)(0, 0, step, 4)
    key = (C.c_uint32 * 2)(seed & 0xffffffff, seed >> 32)
    out = (C.c_uint32 * 4)()
    ob.lib().orc_philox4x32_10(ctr, key, out)
    return _u53(out[0], out[1])


def _make_weights(n, seed):
    rng = np.random.default_rng(seed)
    w = rng.lognormal(0.0, 2.0, n)
    w[rng.random(n) < 0.25] = 0.0
    w[: n // 3] *= 1e-6          # most of the mass sits in the later shards
    return w / w.sum()


def _worker(rank, world, port, n, seed, errors):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from quickmcl_b200 import sharded
        comm = sharded.TorchComm(device="cpu")
        st = comm.struct

        # ---- 1. the callback table, through the C function pointers
        buf = np.zeros(world * 3, np.float64)
        buf[rank * 3:(rank + 1) * 3] = [rank + 0.25, rank + 0.5, rank + 0.75]
        assert st.allgather(None, buf.ctypes.data, 24, None) == 0
        want = np.concatenate([[r + 0.25, r + 0.5, r + 0.75] for r in range(world)])
        assert np.array_equal(buf, want)
        for dt, code in ((np.int32, 0), (np.int64, 1), (np.float64, 2)):
            a = (np.arange(5) + rank * 10).astype(dt)
            assert st.allreduce(None, a.ctypes.data, 5, code, 0, None) == 0      # sum
            assert np.array_equal(a, sum((np.arange(5) + r * 10) for r in range(world)).astype(dt))
            b = (np.arange(5) * (1 if rank else -1)).astype(dt)
            c = b.copy()
            assert st.allreduce(None, b.ctypes.data, 5, code, 1, None) == 0      # min
            assert st.allreduce(None, c.ctypes.data, 5, code, 2, None) == 0      # max
            assert np.array_equal(b, -np.arange(5).astype(dt))
            assert np.array_equal(c, np.arange(5).astype(dt))

        # bin occupancy: 0/1 bytes summed four at a time as int32 words (cluster.cu, build_bins)
        occ = np.zeros(12, np.uint8)
        occ[rank::3] = 1
        assert st.allreduce(None, occ.ctypes.data, 3, 0, 0, None) == 0
        want_occ = np.zeros(12, np.uint8)
        for r in range(world):
            want_occ[r::3] += 1
        assert np.array_equal(occ, want_occ) and np.array_equal(occ != 0, want_occ != 0)

        # ---- 2. sharded low-variance resampling against the oracle
        from oracle import binding as ob
        w = _make_weights(n, seed)
        first, count = sharded.shard_slice(n, rank, world)
        assert (first, count) == (min(n, rank * -(-n // world)),
                                  min(n, (rank + 1) * -(-n // world)) - min(n, rank * -(-n // world)))
        mine = w[first:first + count]
        # carry chain: one allgather round per shard (resample.cu, step 2)
        carry = [0.0] * (world + 1)
        cdf = None
        for r in range(world):
            slot = np.zeros(world)
            if r == rank:
                cdf = np.cumsum(np.concatenate([[carry[r]], mine]))[1:]   # sequential adds
                slot[rank] = cdf[-1] if count else carry[r]
            assert st.allgather(None, slot.ctypes.data, 8, None) == 0
            carry[r + 1] = slot[r]
        M = n
        minv = float(np.float32(1.0) / np.float32(M))
        r0 = _lv_offset(3, 7) * minv
        lo, hi = sharded.lv_owner_range(M, r0, minv, carry[rank], carry[rank + 1],
                                        rank == 0, rank == world - 1)
        m = np.arange(lo, hi, dtype=np.float64)
        U = r0 + m * minv                      # numpy: a rounded product, then a rounded sum
        local = np.minimum(np.searchsorted(cdf, U, side="left"), count - 1)
        src = (local + first).astype(np.int64)  # global source index of each owned slot
        # rebalance to equal blocks with the comm's alltoallv (resample.cu, step 4)
        ranges = np.zeros(2 * world, np.int64)
        ranges[2 * rank:2 * rank + 2] = [lo, hi]
        assert st.allgather(None, ranges.ctypes.data, 16, None) == 0
        m_lo, m_hi = ranges[0::2], ranges[1::2]
        assert m_lo[0] == 0 and m_hi[-1] == M and np.array_equal(m_lo[1:], m_hi[:-1])
        my_a, my_c = sharded.shard_slice(M, rank, world)
        sb = (C.c_uint64 * world)()
        so = (C.c_uint64 * world)()
        rb = (C.c_uint64 * world)()
        ro = (C.c_uint64 * world)()
        for d in range(world):
            a, c = sharded.shard_slice(M, d, world)
            s_lo, s_hi = max(lo, a), min(hi, a + c)
            sb[d] = max(0, s_hi - s_lo) * 8
            so[d] = (s_lo - lo) * 8 if s_hi > s_lo else 0
            r_lo, r_hi = max(int(m_lo[d]), my_a), min(int(m_hi[d]), my_a + my_c)
            rb[d] = max(0, r_hi - r_lo) * 8
            ro[d] = (r_lo - my_a) * 8 if r_hi > r_lo else 0
        recv = np.full(my_c, -1, np.int64)
        assert st.alltoallv(None, src.ctypes.data if len(src) else None, sb, so,
                            recv.ctypes.data, rb, ro, None) == 0
        # the oracle on the whole set
        occ = np.zeros((20, 20), np.int8)
        om = ob.OracleMap()
        om.new_map(occ, 0.05, (0.0, 0.0), field_mode=ob.FIELD_EXACT_EDT)
        of = ob.OracleFilter(om, seed=3)
        of.set_parameters(resample_type=0, particle_count_min=n, particle_count_max=n)
        rng = np.random.default_rng(1)
        of.set_particles(ob.make_particles(rng.random(n), rng.random(n), rng.random(n), w))
        of.set_rng(3, 7)
        of.resample_and_cluster()
        idx = of.resample_indices()
        assert np.array_equal(recv, idx[my_a:my_a + my_c]), \
            f"rank {rank}: {np.flatnonzero(recv != idx[my_a:my_a + my_c])[:5]}"
        assert comm.calls["alltoallv"] == 1 and comm.error is None
        dist.barrier()
        dist.destroy_process_group()
    except Exception:  # noqa: BLE001
        import traceback
        errors.put((rank, traceback.format_exc()))
        raise


@pytest.mark.parametrize("n,seed", [(5000, 1), (4097, 2), (10, 3)])
def test_sharded_low_variance_protocol_gloo_world2(n, seed):
    world = 2
    ctx = mp.get_context("spawn")
    errors = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, errors))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            msgs.append((-1, "worker timed out"))
    assert not msgs, "\n".join(f"rank {r}:\n{m}" for r, m in msgs)
    assert all(p.exitcode == 0 for p in procs)


def test_shard_slice_and_owner_range_tile_everything():
    """Pure host arithmetic (no process group): blocks tile [0, n) and owner ranges
    tile [0, M) for any split of a CDF."""
    from quickmcl_b200 import sharded
    for n in (0, 1, 7, 8, 9, 1000, 16_000_000):
        for size in (1, 2, 3, 8):
            at = 0
            for r in range(size):
                a, c = sharded.shard_slice(n, r, size)
                assert a == at
                at += c
            assert at == n
    rng = np.random.default_rng(5)
    M = 100_003
    w = rng.random(M)
    w[rng.random(M) < 0.3] = 0.0
    w /= w.sum() * 1.0000001          # total a hair below 1: the clamped overrun is exercised
    cdf = np.cumsum(w)
    minv = float(np.float32(1.0) / np.float32(M))
    r0 = 0.37 * minv
    U = r0 + np.arange(M, dtype=np.float64) * minv
    want = np.minimum(np.searchsorted(cdf, U, side="left"), M - 1)   # the reference's index
    for size in (2, 3, 8):
        cuts = [sharded.shard_slice(M, r, size)[0] for r in range(size)] + [M]
        at = 0
        for r in range(size):
            c_lo = cdf[cuts[r] - 1] if cuts[r] else 0.0
            c_hi = cdf[cuts[r + 1] - 1]
            lo, hi = sharded.lv_owner_range(M, r0, minv, c_lo, c_hi, r == 0, r == size - 1)
            assert lo == at
            at = hi
            assert np.all((want[lo:hi] >= cuts[r]) & (want[lo:hi] < cuts[r + 1]))
        assert at == M
