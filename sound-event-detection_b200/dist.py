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
    """Multi-GPU assembly of `framewise_output` / `clipwise_output` on rank `dst` WITHOUT a gather pass and without a
    collective: the destination GPU owns one buffer for all ranks' results (`slots` of them, used round-robin), every
    other rank maps it through CUDA IPC (sed_peer_* of the C ABI) and delivers its slice either as the pooling-head
    kernels' own stores over NVLink / NVSwitch (`PackedModel.forward(out=peer.local_out(step))`) or by DMA
    (`push`, on a side stream: the transfer overlaps the next batch's kernels without occupying an SM).

    Completion is flag based: behind its data a rank copies the word `step + 1` into its entry of the destination's
    flag array (`signal`); `complete(step)` makes the destination's stream wait until every entry has reached that
    value (driver stream memory operations: no kernel, no SM).  The ranks never meet in a collective, so each runs at
    its own pace -- a per-step all-reduce makes every step as slow as the slowest of N GPUs (measured at N = 8:
    22.0 ms per step against 21.3 ms without any assembly).  Back-pressure: when the destination calls
    `complete(k)` it releases the slots of steps < k to the producers (`ack` words in their memory); a producer's
    transfer of step j waits for the release of step j - slots.  So the results of step k stay valid on the
    destination until it calls `complete(k + 1)`, and a producer runs at most `slots` steps ahead.
    Where stream memory operations are unavailable, completion falls back to one 4-byte all-reduce per step.

    Replaces the gather-to-GPU-0 of torch.nn.DataParallel (pytorch/main_strong.py:541) for one-process-per-GPU runs."""

    def __init__(self, clips_per_rank, frames, classes, device, dst=0, group=None, slots=4, use_flags=True):
        from . import capi
        import ctypes
        self._ct = ctypes
        self.group, self.dst, self.slots = group, dst, int(slots)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n, self.frames, self.classes = int(clips_per_rank), int(frames), int(classes)
        self.device = torch.device(device)
        lib = capi.load()
        total = self.world * self.n
        self._frame_elems, self._clip_elems = total * self.frames * self.classes, total * self.classes
        slot_bytes = (4 * (self._frame_elems + self._clip_elems) + 255) // 256 * 256
        self._slot_bytes = slot_bytes
        self._flags_off = slot_bytes * self.slots          # destination: one 4-byte flag per source rank
        self._lib, self._base, self._owner = lib, ctypes.c_void_p(), self.rank == dst
        self._small = ctypes.c_void_p()                     # every rank: [ack word | staging word | staging word 2]
        self._peer_small = {}                               # destination: rank -> mapped pointer of that rank's words
        self._opened = []
        self._closed = False
        err = ""
        with torch.cuda.device(self.device):
            hbig, hsmall = (ctypes.c_ubyte * 64)(), (ctypes.c_ubyte * 64)()
            try:
                capi.check(lib.sed_peer_alloc(256, ctypes.byref(self._small)), "sed_peer_alloc")
                capi.check(lib.sed_peer_export(self._small, hsmall), "sed_peer_export")
                if self._owner:
                    capi.check(lib.sed_peer_alloc(self._flags_off + 256, ctypes.byref(self._base)), "sed_peer_alloc")
                    capi.check(lib.sed_peer_export(self._base, hbig), "sed_peer_export")
                zero = torch.zeros(64 + (self.world * 4 + 255) // 4, dtype=torch.int32, device=self.device)
                capi.check(lib.sed_peer_copy(self._small, capi.ptr(zero), 256, capi.current_stream(self.device)), "sed_peer_copy")
                if self._owner:
                    capi.check(lib.sed_peer_copy(ctypes.c_void_p(self._base.value + self._flags_off), capi.ptr(zero), 256,
                                                 capi.current_stream(self.device)), "sed_peer_copy")
                torch.cuda.synchronize(self.device)
            except Exception as e:  # noqa: BLE001 -- reported to every rank below
                err = "rank %d: %s" % (self.rank, e)
            # stream memory operations on this device?
            flags_ok = bool(use_flags) and not err
            if flags_ok:
                try:
                    stage = ctypes.c_void_p(self._small.value + 8)
                    capi.check(lib.sed_stream_write32(stage, 7, capi.current_stream(self.device)), "sed_stream_write32")
                    capi.check(lib.sed_stream_wait_geq32(stage, 7, capi.current_stream(self.device)), "sed_stream_wait_geq32")
                    torch.cuda.synchronize(self.device)
                except Exception:  # noqa: BLE001
                    flags_ok = False
            infos = [None] * self.world
            dist.all_gather_object(infos, (bytes(hbig), bytes(hsmall), err, flags_ok), group=group)
            errs = [i[2] for i in infos if i[2]]
            if not errs:
                try:
                    if self._owner:
                        for r, info in enumerate(infos):
                            if r != dst:
                                p = ctypes.c_void_p()
                                h = (ctypes.c_ubyte * 64).from_buffer_copy(info[1])
                                capi.check(lib.sed_peer_open(h, ctypes.byref(p)), "sed_peer_open")
                                self._peer_small[r] = p
                                self._opened.append(p)
                    else:
                        h = (ctypes.c_ubyte * 64).from_buffer_copy(infos[dst][0])
                        capi.check(lib.sed_peer_open(h, ctypes.byref(self._base)), "sed_peer_open")
                        self._opened.append(self._base)
                except Exception as e:  # noqa: BLE001
                    err = "rank %d: %s" % (self.rank, e)
            errs2 = [None] * self.world
            dist.all_gather_object(errs2, err, group=group)
            errs += [e for e in errs2 if e and e not in errs]
            if errs:
                self.close()
                raise RuntimeError("peer-memory result buffers unavailable: " + "; ".join(errs))
        self.use_flags = all(i[3] for i in infos)
        self._token = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._released = 0
        self._ack_stream = torch.cuda.Stream(self.device) if (self.use_flags and self._owner) else None

    # ------------------------------------------------------------------ addressing
    def _views(self, slot):
        base = self._base.value + (slot % self.slots) * self._slot_bytes
        total = self.world * self.n
        frame = _PeerView(base, (total, self.frames, self.classes))
        clip = _PeerView(base + 4 * self._frame_elems, (total, self.classes))
        return clip, frame

    def local_out(self, step):
        """(clipwise, framewise) output buffers of THIS rank for step `step`: its [n, ...] slices of the destination's
        buffer.  Pass as `PackedModel.forward(wave, out=...)`, then call `signal(step)` on the same stream."""
        clip, frame = self._views(step)
        lo, hi = self.rank * self.n, (self.rank + 1) * self.n
        return clip[lo:hi], frame[lo:hi]

    def _stream(self):
        from . import capi
        return capi.current_stream(self.device)

    # ------------------------------------------------------------------ producer side
    def wait_turn(self, step):
        """Current stream: wait until the destination has released the slot step `step` is going to overwrite."""
        from . import capi
        if self.use_flags and not self._owner and step >= self.slots:
            capi.check(self._lib.sed_stream_wait_geq32(self._small, step - self.slots + 1, self._stream()),
                       "sed_stream_wait_geq32")

    def push(self, clipwise, framewise, step):
        """Copy-engine variant: this rank's finished [n, ...] results (ordinary tensors on its own GPU) go to its slice
        of the destination's buffer by DMA on the CURRENT stream (sed_peer_copy), followed by the completion flag."""
        from . import capi
        self.wait_turn(step)
        clip, frame = self.local_out(step)
        stream = self._stream()
        for src, dst in ((clipwise, clip), (framewise, frame)):
            if not src.is_contiguous() or src.dtype != torch.float32 or tuple(src.shape) != dst.shape:
                raise ValueError("push expects contiguous float32 results of shape %s" % (dst.shape,))
            capi.check(self._lib.sed_peer_copy(dst.data_ptr(), src.data_ptr(), src.numel() * 4, stream), "sed_peer_copy")
        self.signal(step)

    def signal(self, step):
        """Current stream: everything this rank queued so far for step `step` is in the destination's memory."""
        from . import capi
        if not self.use_flags:
            return
        ct = self._ct
        stream = self._stream()
        flag = ct.c_void_p(self._base.value + self._flags_off + 4 * self.rank)
        if self._owner:
            capi.check(self._lib.sed_stream_write32(flag, (step + 1) & 0xFFFFFFFF, stream), "sed_stream_write32")
        else:
            stage = ct.c_void_p(self._small.value + 8)
            capi.check(self._lib.sed_stream_write32(stage, (step + 1) & 0xFFFFFFFF, stream), "sed_stream_write32")
            capi.check(self._lib.sed_peer_copy(flag, stage, 4, stream), "sed_peer_copy")

    # ------------------------------------------------------------------ consumer side
    def complete(self, step):
        """Stream-ordered completion of step `step`.  On `dst`: the current stream waits for every rank's flag, the
        slots of all earlier steps are released to the producers, and the assembled dict is returned (torch tensors
        aliasing the buffer, valid until `complete(step + 1)`); elsewhere a no-op returning None."""
        from . import capi
        if not self.use_flags:
            dist.all_reduce(self._token, group=self.group)
        elif self._owner:
            ct = self._ct
            stream = self._stream()
            for r in range(self.world):
                flag = ct.c_void_p(self._base.value + self._flags_off + 4 * r)
                capi.check(self._lib.sed_stream_wait_geq32(flag, (step + 1) & 0xFFFFFFFF, stream), "sed_stream_wait_geq32")
            if step > self._released:  # release the slots of steps < step (on a side stream: seven 4-byte copies)
                here = torch.cuda.Event()
                here.record(torch.cuda.current_stream(self.device))
                self._ack_stream.wait_event(here)
                ack = ct.c_void_p(self._ack_stream.cuda_stream)
                stage = ct.c_void_p(self._small.value + 16)
                capi.check(self._lib.sed_stream_write32(stage, step & 0xFFFFFFFF, ack), "sed_stream_write32")
                for r, p in self._peer_small.items():
                    capi.check(self._lib.sed_peer_copy(p, stage, 4, ack), "sed_peer_copy")
                self._released = step
        if not self._owner:
            return None
        clip, frame = self._views(step)
        return {"framewise_output": torch.as_tensor(frame, device=self.device),
                "clipwise_output": torch.as_tensor(clip, device=self.device)}

    def close(self):
        if self._closed:
            return
        self._closed = True
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                if p.value:
                    self._lib.sed_peer_close(p)
            if self._owner and self._base.value:
                self._lib.sed_peer_free(self._base)
            if self._small.value:
                self._lib.sed_peer_free(self._small)
        self._base.value = None
        self._small.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
