"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard maths and the visibility
all_gather.  The per-rank compute is the oracle here -- the CUDA path is covered by the GPU
tests; what this checks is that sharding + gathering reproduces the single-rank result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gr_clenabled_b200 import multigpu
from oracle import oracle as orc


def test_shard_range_covers_everything():
    for n in (1, 7, 64, 1000, 1024):
        for world in (1, 2, 3, 8):
            parts = [multigpu.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts[:-1], parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, F, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        A, T, npol = 5, 48, 2
        buf = orc.rng_i8(T * A * F * npol * 2, orc.SEED_X).reshape(T, A, F, npol, 2)
        first, count = multigpu.shard_channels(F, rank, world)
        slab = np.ascontiguousarray(buf[:, :, first:first + count])
        local = orc.xengine_exact(slab, A, count, T, npol)                 # [count*nbl*npol^2, 2]
        nbl = A * (A + 1) // 2
        full = multigpu.gather_visibilities(torch.from_numpy(local), F, nbl * npol * npol)
        want = orc.xengine_exact(buf, A, F, T, npol)
        ok = np.array_equal(full.numpy().reshape(-1, 2), want)
        # FFT vectors: every rank transforms its range, concatenation equals the whole
        nvec, N = 9, 64
        x = orc.rng_c32(nvec * N, orc.SEED_F)
        vf, vc = multigpu.shard_vectors(nvec, rank, world)
        mine = torch.from_numpy(orc.fft(x[vf * N:(vf + vc) * N], N, -1).view(np.float32).copy())
        sizes = [multigpu.shard_vectors(nvec, r, world)[1] * N * 2 for r in range(world)]
        bufs = [torch.empty(max(sizes)) for _ in range(world)]
        pad = torch.zeros(max(sizes))
        pad[:mine.numel()] = mine
        dist.all_gather(bufs, pad)
        whole = torch.cat([b[:s] for b, s in zip(bufs, sizes)]).numpy().view(np.complex64)
        ok = ok and np.array_equal(whole, orc.fft(x, N, -1))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("F", [16, 7])          # even and ragged channel split
def test_sharded_xengine_and_fft_match_single_rank(F):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, F, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    res = dict(q.get(timeout=5) for _ in procs)
    assert res == {0: True, 1: True}
