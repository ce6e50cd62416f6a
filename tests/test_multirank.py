"""CPU tier: the N>1 path (batch sharding, timing reduction, optional gather) with world_size 2
over gloo.  The per-shard computation uses the oracle's CPU port, which stands in for the GPU
kernels here; what is under test is that sharding needs no data-path collective."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from kapre_b200.sharding import gather_shards, reduce_max, shard_range
        from oracle.fast_cpu import MelSpectrogramCPU
        torch.set_num_threads(1)
        B = 6
        g = torch.Generator().manual_seed(7)
        x = torch.rand((B, 6000, 1), generator=g) * 2 - 1
        x[1] *= 1e-3   # an item whose dB clamp differs from its neighbours'
        model = MelSpectrogramCPU(n_fft=512, hop_length=128, sample_rate=16000, n_mels=40, return_decibel=True,
                                  db_dynamic_range=30.0)
        lo, hi = shard_range(B, rank, world)
        y_local = model(x[lo:hi])
        y_all = gather_shards(y_local)
        full = model(x)
        ok = bool(torch.equal(y_all, full))
        tmax = reduce_max(float(rank + 1))
        if rank == 0:
            ret.put((ok, tmax, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_world2_sharding_gloo():
    ctx = mp.get_context('spawn')
    ret = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, tmax, rng = ret.get()
    assert ok, 'concatenated shard outputs differ from the unsharded output'
    assert tmax == 2.0
    assert rng == (0, 3)


def test_shard_range_covers_batch():
    from kapre_b200.sharding import shard_range
    for n in (0, 1, 7, 256, 8192):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
