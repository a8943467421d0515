"""Multi-GPU plumbing: one process per GPU, contiguous batch shards, no collective on the hot path.

The reference's only multi-GPU mechanism is `torch.nn.DataParallel` (scatter on dim 0, replicate the
module every call, gather to GPU 0; call sites pytorch/main_strong.py:541, pytorch/predict.py:239).
Clips are independent (eval-mode BatchNorm, per-clip GRU/attention, top_db=None), so each rank runs the
whole path on its shard and a single gather assembles `framewise_output` / `clipwise_output` on rank 0
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def bind_host_to_gpu(device_index):
    """Pin the calling process to the CPUs (NUMA node) nearest to GPU `device_index` (NVML's ideal affinity), so that
    the pinned staging buffers it allocates afterwards are local to that GPU's PCIe root.  With eight ranks copying
    5 GB per step, buffers on the wrong socket cross the inter-socket link.  Returns a short description; never raises
    (containers may forbid affinity changes) and keeps the old mask if NVML offers fewer than two CPUs."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        before = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        after = os.sched_getaffinity(0)
        if len(after) < 2:
            os.sched_setaffinity(0, before)
            return "kept (%d cpus; NVML offered %d)" % (len(before), len(after))
        return "bound to %d of %d cpus" % (len(after), len(before))
    except Exception as e:  # noqa: BLE001 -- best effort by design
        return "unavailable (%s)" % (str(e)[:80] or type(e).__name__)


def shard_bounds(batch, world_size, rank):
    """Contiguous split of `batch` clips: rank r owns [lo, hi). Remainder goes to the first ranks."""
    base, rem = divmod(batch, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_batch(wave, world_size=None, rank=None):
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(wave.shape[0], world_size, rank)
    return wave[lo:hi]


def gather_outputs(out, dst=0, group=None, shard_sizes=None):
    """Gather the per-rank output dict to rank `dst` (concatenated along the batch dim).

    Returns the assembled dict on `dst`, None elsewhere.  Keys gathered: framewise_output,
    clipwise_output (the two tensors every reference caller reads, pytorch_utils.py:57-62).
    Ragged shards are supported when `shard_sizes` (list of per-rank batch sizes) is given.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return {k: out[k] for k in ("framewise_output", "clipwise_output")}
    result = {}
    for key in ("framewise_output", "clipwise_output"):
        t = out[key].contiguous()
        if shard_sizes is None:
            sizes = [t.shape[0]] * world
        else:
            sizes = list(shard_sizes)
        if rank == dst:
            bufs = [torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for n in sizes]
            if len(set(sizes)) == 1:
                dist.gather(t, gather_list=bufs, dst=dst, group=group)
            else:
                _ragged_gather(t, bufs, dst, group, rank, world)
            result[key] = torch.cat(bufs, 0)
        else:
            if len(set(sizes)) == 1:
                dist.gather(t, gather_list=None, dst=dst, group=group)
            else:
                _ragged_gather(t, None, dst, group, rank, world)
    return result if rank == dst else None


def _ragged_gather(t, bufs, dst, group, rank, world):
    if rank == dst:
        reqs = []
        for r in range(world):
            if r == dst:
                bufs[r].copy_(t)
            else:
                reqs.append(dist.irecv(bufs[r], src=r, group=group))
        for q in reqs:
            q.wait()
    else:
        dist.send(t, dst=dst, group=group)
