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


def gather_outputs(out, dst=0, group=None, shard_sizes=None, into=None, async_op=False):
    """Gather the per-rank output dict to rank `dst` (concatenated along the batch dim) with the process group's
    backend (NCCL send/recv over NVLink on GPUs, gloo in the CPU tests).

    Returns the assembled dict on `dst`, None elsewhere.  Keys gathered: framewise_output,
    clipwise_output (the two tensors every reference caller reads, pytorch_utils.py:57-62).
    Ragged shards are supported when `shard_sizes` (list of per-rank batch sizes) is given.
    into: optional dict of preallocated [sum(sizes), ...] tensors on `dst` -- every shard is received straight
    into its slice (no concatenation pass over the assembled result).
    async_op: return (result, works) without waiting; the caller waits on the works (equal shards only)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return {k: out[k] for k in ("framewise_output", "clipwise_output")}
    result, works = {}, []
    for key in ("framewise_output", "clipwise_output"):
        t = out[key].contiguous()
        sizes = [t.shape[0]] * world if shard_sizes is None else list(shard_sizes)
        equal = len(set(sizes)) == 1
        if rank == dst:
            full = into[key] if into is not None else torch.empty((sum(sizes),) + tuple(t.shape[1:]), dtype=t.dtype,
                                                                 device=t.device)
            bounds = [sum(sizes[:r]) for r in range(world + 1)]
            bufs = [full[bounds[r]:bounds[r + 1]] for r in range(world)]  # contiguous slices: received in place
            if equal:
                w = dist.gather(t, gather_list=bufs, dst=dst, group=group, async_op=async_op)
            else:
                w = _ragged_gather(t, bufs, dst, group, rank, world)
            result[key] = full
        else:
            if equal:
                w = dist.gather(t, gather_list=None, dst=dst, group=group, async_op=async_op)
            else:
                w = _ragged_gather(t, None, dst, group, rank, world)
        works.append(w)
    res = result if rank == dst else None
    return (res, [w for w in works if w is not None]) if async_op else res


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


class _PeerView:
    """A float32 array in peer-mapped device memory: what the engine needs of an output buffer (`data_ptr`, shape,
    slicing along dim 0).  The memory belongs to another process's GPU; kernels of this rank store into it over NVLink."""

    def __init__(self, ptr, shape):
        self._ptr, self.shape = int(ptr), tuple(shape)

    def data_ptr(self):
        return self._ptr

    def __getitem__(self, sl):
        if not isinstance(sl, slice) or sl.step not in (None, 1):
            raise IndexError("peer views slice along dim 0 only")
        lo, hi, _ = sl.indices(self.shape[0])
        row = 4
        for d in self.shape[1:]:
            row *= d
        return _PeerView(self._ptr + lo * row, (max(0, hi - lo),) + self.shape[1:])

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f4", "data": (self._ptr, False), "version": 2}


class PeerGather:
    """Multi-GPU assembly of `framewise_output` / `clipwise_output` on rank `dst` WITHOUT a gather pass: the
    destination GPU owns one buffer for all ranks' results (`slots` of them, used round-robin), every other rank maps
    it through CUDA IPC (sed_peer_* of the C ABI) and hands its slice to the pooling-head kernels as their output
    pointers (`PackedModel.forward(out=...)`), so the results leave each GPU as the kernel's own stores over
    NVLink / NVSwitch.  `complete()` is one stream-ordered 4-byte all-reduce: once the destination's stream has
    passed it, every rank's head kernel has finished and the assembled tensors are readable there.

    Replaces the gather-to-GPU-0 of torch.nn.DataParallel (pytorch/main_strong.py:541) for one-process-per-GPU runs.
    Protocol: step k writes slot k % slots; the destination must consume slot k on the stream that later calls
    complete() for step k + slots - 1 or earlier (with slots=2: before its next complete())."""

    def __init__(self, clips_per_rank, frames, classes, device, dst=0, group=None, slots=2):
        from . import capi
        import ctypes
        self.group, self.dst, self.slots = group, dst, int(slots)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n, self.frames, self.classes = int(clips_per_rank), int(frames), int(classes)
        self.device = torch.device(device)
        lib = capi.load()
        total = self.world * self.n
        self._frame_elems, self._clip_elems = total * self.frames * self.classes, total * self.classes
        slot_bytes = 4 * (self._frame_elems + self._clip_elems)
        slot_bytes = (slot_bytes + 255) // 256 * 256
        self._slot_bytes = slot_bytes
        self._lib, self._base, self._owner = lib, ctypes.c_void_p(), self.rank == dst
        handle = (ctypes.c_ubyte * 64)()
        err = ""
        with torch.cuda.device(self.device):
            try:
                if self._owner:
                    capi.check(lib.sed_peer_alloc(slot_bytes * self.slots, ctypes.byref(self._base)), "sed_peer_alloc")
                    capi.check(lib.sed_peer_export(self._base, handle), "sed_peer_export")
            except Exception as e:  # noqa: BLE001 -- reported to every rank below
                err = str(e)
            box = [(bytes(handle), err)]
            dist.broadcast_object_list(box, src=dst, group=group)
            raw, err = box[0]
            if not err and not self._owner:
                try:
                    h = (ctypes.c_ubyte * 64).from_buffer_copy(raw)
                    capi.check(lib.sed_peer_open(h, ctypes.byref(self._base)), "sed_peer_open")
                except Exception as e:  # noqa: BLE001
                    err = "rank %d: %s" % (self.rank, e)
            errs = [None] * self.world
            dist.all_gather_object(errs, err, group=group)
            errs = [e for e in errs if e]
            if errs:
                self.close()
                raise RuntimeError("peer-memory result buffers unavailable: " + "; ".join(errs))
        self._token = torch.zeros(1, dtype=torch.int32, device=self.device)

    def _views(self, slot):
        base = self._base.value + (slot % self.slots) * self._slot_bytes
        total = self.world * self.n
        frame = _PeerView(base, (total, self.frames, self.classes))
        clip = _PeerView(base + 4 * self._frame_elems, (total, self.classes))
        return clip, frame

    def local_out(self, slot):
        """(clipwise, framewise) output buffers of THIS rank for step `slot`: its [n, ...] slices of the destination's
        buffer.  Pass as `PackedModel.forward(wave, out=...)`."""
        clip, frame = self._views(slot)
        lo, hi = self.rank * self.n, (self.rank + 1) * self.n
        return clip[lo:hi], frame[lo:hi]

    def push(self, clipwise, framewise, slot):
        """Copy-engine variant: this rank's finished [n, ...] results (ordinary tensors on its own GPU) go to its slice
        of the destination's buffer by DMA on the CURRENT stream (sed_peer_copy) -- run it on a side stream and the
        transfer overlaps the next batch's kernels without occupying an SM."""
        from . import capi
        clip, frame = self.local_out(slot)
        stream = capi.current_stream(self.device)
        for src, dst in ((clipwise, clip), (framewise, frame)):
            if not src.is_contiguous() or src.dtype != torch.float32 or tuple(src.shape) != dst.shape:
                raise ValueError("push expects contiguous float32 results of shape %s" % (dst.shape,))
            capi.check(self._lib.sed_peer_copy(dst.data_ptr(), src.data_ptr(), src.numel() * 4, stream), "sed_peer_copy")

    def complete(self, slot):
        """Stream-ordered completion of step `slot` on every rank.  Returns the assembled dict on `dst` (torch tensors
        aliasing the buffer, valid until the slot is written again), None elsewhere."""
        dist.all_reduce(self._token, group=self.group)
        if not self._owner:
            return None
        clip, frame = self._views(slot)
        return {"framewise_output": torch.as_tensor(frame, device=self.device),
                "clipwise_output": torch.as_tensor(clip, device=self.device)}

    def close(self):
        if getattr(self, "_base", None) is not None and self._base.value:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                (self._lib.sed_peer_free if self._owner else self._lib.sed_peer_close)(self._base)
            self._base.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
