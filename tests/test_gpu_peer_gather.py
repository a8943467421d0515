"""Two-rank NCCL run on two GPUs: the pooling-head kernels of rank 1 store their outputs straight into rank 0's
buffer (dist.PeerGather, CUDA-IPC peer memory) and the assembled result equals one GPU running the whole batch,
bit for bit; the in-place NCCL gather gives the same."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from sed_b200 import dist as sdist
    from sed_b200 import engine, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    mt = "Cnn_9layers_Gru_FrameAtt"
    pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
    n = 5
    full = synth.synthetic_waveform(world * n, 48000, seed=61, kind="events")
    mine = full[rank * n:(rank + 1) * n].to(dev)
    frames = pm.frames_for((48000 // 160 + 1) // 8)
    ok = True
    for use_flags in (True, False):  # flag-based completion, then the all-reduce fallback
        peer = sdist.PeerGather(n, frames, pm.classes, dev, dst=0, slots=3, use_flags=use_flags)
        if use_flags and not peer.use_flags:
            print("stream memory operations unavailable on this box: flags not exercised")
        for step in range(7):  # slots are reused round-robin; even steps: fused peer stores, odd steps: DMA push
            if step % 2 == 0:
                peer.wait_turn(step)
                pm.forward(mine, out=peer.local_out(step))
                peer.signal(step)
            else:
                loc = pm.forward(mine)
                peer.push(loc["clipwise_output"], loc["framewise_output"], step)
            res = peer.complete(step)
            if rank == 0:
                ref = pm.forward(full.to(dev))
                torch.cuda.synchronize()
                ok = ok and torch.equal(res["framewise_output"], ref["framewise_output"])
                ok = ok and torch.equal(res["clipwise_output"], ref["clipwise_output"])
            else:
                assert res is None
        torch.cuda.synchronize()
        dist.barrier()
        peer.close()
    out = pm.forward(mine)
    into = {"framewise_output": torch.empty((world * n, frames, pm.classes), device=dev),
            "clipwise_output": torch.empty((world * n, pm.classes), device=dev)} if rank == 0 else None
    got = sdist.gather_outputs(out, dst=0, into=into)
    if rank == 0:
        torch.cuda.synchronize()
        ok = ok and torch.equal(got["framewise_output"], ref["framewise_output"])
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_gather_two_gpus():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ok
