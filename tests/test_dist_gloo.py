"""N > 1 host path on CPU: two gloo ranks shard a batch and gather outputs to rank 0 (no data-path collective)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sed_b200 import dist as sdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_forward(wave):
    # per-clip function (stand-in for the GPU path): outputs depend only on the clip itself
    B = wave.shape[0]
    clip = torch.sigmoid(wave[:, :25])
    frame = torch.sigmoid(wave[:, :50].reshape(B, 2, 25)).repeat(1, 4, 1)
    return {"clipwise_output": clip, "framewise_output": frame}


def _worker(rank, world, port, batch, ragged, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    wave = torch.randn(batch, 64, generator=g)
    shard = sdist.shard_batch(wave)
    out = _fake_forward(shard)
    sizes = [hi - lo for lo, hi in (sdist.shard_bounds(batch, world, r) for r in range(world))]
    res = sdist.gather_outputs(out, dst=0, shard_sizes=sizes if ragged else None)
    # the same gather received in place into a preallocated result (no concatenation pass)
    into = {k: torch.full((batch,) + tuple(v.shape[1:]), -1.0) for k, v in out.items()} if rank == 0 else None
    res2 = sdist.gather_outputs(out, dst=0, shard_sizes=sizes if ragged else None, into=into)
    if rank == 0:
        full = _fake_forward(wave)
        ok = all(torch.equal(res[k], full[k]) for k in full)
        ok = ok and all(res2[k] is into[k] and torch.equal(into[k], full[k]) for k in full)
        q.put(ok)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def _run(batch, ragged):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_two_rank_gather_even():
    _run(8, False)


def test_two_rank_gather_ragged():
    _run(7, True)
