"""Platform copy ceiling of the end-to-end step: what the host <-> device links of this box deliver when every rank
does NOTHING but the copies of one bench step (no kernels).

  torchrun --nproc-per-node N tools/h2d_ceiling.py [--clips 1024] [--steps 8] > profiles/r02_h2d_ceiling.json

Per step and rank the headline workload (bench.py, 1024 clips of 10 s at 16 kHz) moves
  H2D  655.36 MB of float32 waveform  (327.68 MB as int16 PCM, the reference's storage format utilities.py:78-79)
  D2H  102.50 MB of framewise + clipwise float32 results.
Scenarios (all ranks concurrently unless marked solo; one cudaMemcpyAsync per span, pinned host memory):
  h2d_f32 / h2d_i16 / d2h             one direction alone
  step_f32 / step_i16                 both directions on two streams (what a fully pipelined step needs)
  step_*_solo                         rank 0 alone (single-link ceiling)
  h2d_f32_chunk32 / chunk4            the same bytes as 32 MB / 4 MB spans
  h2d_f32_wc                          write-combined pinned source (cudaHostAllocWriteCombined)
  h2d_f32_thp                         transparent-huge-page source registered with cudaHostRegister
The implied step ceiling in clips/s = N * clips / (max-over-ranks step time).  Rank 0 prints one JSON object.
"""
import argparse
import ctypes
import json
import mmap
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def cudart():
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    for name in sorted(os.listdir(libdir)):
        if name.startswith("libcudart"):
            return ctypes.CDLL(os.path.join(libdir, name))
    import glob
    for pat in ("/usr/local/cuda/lib64/libcudart.so*",
                os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        hits = sorted(glob.glob(pat))
        if hits:
            return ctypes.CDLL(hits[0])
    raise OSError("libcudart not found")


def wc_pinned(rt, nbytes):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))
    if rc != 0:
        raise RuntimeError("cudaHostAlloc(WC) -> %d" % rc)
    return p


def thp_pinned(rt, nbytes):
    """anonymous mapping, 2 MB aligned, MADV_HUGEPAGE, touched, then registered."""
    size = (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
    m = mmap.mmap(-1, size + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = ctypes.addressof(ctypes.c_char.from_buffer(m))
    aligned = (base + (2 << 20) - 1) // (2 << 20) * (2 << 20)
    libc = ctypes.CDLL(None, use_errno=True)
    libc.madvise(ctypes.c_void_p(aligned), ctypes.c_size_t(size), 14)  # MADV_HUGEPAGE
    ctypes.memset(aligned, 1, size)
    rc = rt.cudaHostRegister(ctypes.c_void_p(aligned), ctypes.c_size_t(size), ctypes.c_uint(0))
    if rc != 0:
        raise RuntimeError("cudaHostRegister -> %d" % rc)
    thp = None
    try:
        with open("/proc/self/smaps_rollup") as f:
            for ln in f:
                if ln.startswith("AnonHugePages"):
                    thp = ln.split()[1] + " kB"
    except OSError:
        pass
    return m, ctypes.c_void_p(aligned), thp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    rt = cudart()
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]

    n_f32 = args.clips * 160000 * 4
    n_i16 = args.clips * 160000 * 2
    n_out = args.clips * (1000 * 25 + 25) * 4
    src = torch.empty(n_f32, dtype=torch.uint8).pin_memory()
    src.fill_(3)
    dst_host = torch.empty(n_out, dtype=torch.uint8).pin_memory()
    dst_host.zero_()
    d_in = torch.empty(n_f32, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n_out, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def h2d(nbytes, chunk=None, host_ptr=None):
        hp = src.data_ptr() if host_ptr is None else host_ptr
        chunk = nbytes if chunk is None else chunk
        for o in range(0, nbytes, chunk):
            n = min(chunk, nbytes - o)
            rc = rt.cudaMemcpyAsync(d_in.data_ptr() + o, hp + o, n, 1, s_in.cuda_stream)
            assert rc == 0, rc

    def d2h():
        rc = rt.cudaMemcpyAsync(dst_host.data_ptr(), d_out.data_ptr(), n_out, 2, s_out.cuda_stream)
        assert rc == 0, rc

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def run(name, fn, nbytes_in, nbytes_out, solo=False):
        active = (rank == 0) or not solo
        if active:
            fn()
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(args.steps):
                fn()
            s_in.synchronize()
            s_out.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        n_active = 1 if solo else world
        step = t.item() / args.steps
        return {"scenario": name, "ranks_active": n_active, "ms_per_step_max_over_ranks": 1e3 * step,
                "h2d_GBps_per_rank": nbytes_in / step / 1e9, "d2h_GBps_per_rank": nbytes_out / step / 1e9,
                "aggregate_GBps": n_active * (nbytes_in + nbytes_out) / step / 1e9,
                "implied_clips_per_s": n_active * args.clips / step}

    rows = []
    rows.append(run("h2d_f32", lambda: h2d(n_f32), n_f32, 0))
    rows.append(run("h2d_i16", lambda: h2d(n_i16), n_i16, 0))
    rows.append(run("d2h", d2h, 0, n_out))
    rows.append(run("step_f32", lambda: (h2d(n_f32), d2h()), n_f32, n_out))
    rows.append(run("step_i16", lambda: (h2d(n_i16), d2h()), n_i16, n_out))
    rows.append(run("step_f32_solo", lambda: (h2d(n_f32), d2h()), n_f32, n_out, solo=True))
    rows.append(run("step_i16_solo", lambda: (h2d(n_i16), d2h()), n_i16, n_out, solo=True))
    rows.append(run("h2d_f32_chunk32", lambda: h2d(n_f32, 32 << 20), n_f32, 0))
    rows.append(run("h2d_f32_chunk4", lambda: h2d(n_f32, 4 << 20), n_f32, 0))
    notes = {}
    try:
        wc = wc_pinned(rt, n_f32)
        ctypes.memset(wc, 5, n_f32)
        rows.append(run("h2d_f32_wc", lambda: h2d(n_f32, host_ptr=wc.value), n_f32, 0))
        rows.append(run("step_f32_wc", lambda: (h2d(n_f32, host_ptr=wc.value), d2h()), n_f32, n_out))
        rt.cudaFreeHost(wc)
    except Exception as e:  # noqa: BLE001
        notes["wc"] = str(e)[:120]
    try:
        keep, hp, thp = thp_pinned(rt, n_f32)
        notes["thp_AnonHugePages"] = thp
        rows.append(run("h2d_f32_thp", lambda: h2d(n_f32, host_ptr=hp.value), n_f32, 0))
        rt.cudaHostUnregister(hp)
    except Exception as e:  # noqa: BLE001
        notes["thp"] = str(e)[:120]
    # host DRAM copy rate of one thread per rank, all ranks at once (what a collate / concatenate costs beside the DMA)
    a = torch.empty(256 << 20, dtype=torch.uint8)
    b = torch.empty(256 << 20, dtype=torch.uint8)
    torch.set_num_threads(1)
    b.copy_(a)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        b.copy_(a)
    dt = (time.perf_counter() - t0) / 4
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    notes["host_memcpy_GBps_per_rank_1thread_read+write"] = 2 * (256 << 20) / t.item() / 1e9
    if rank == 0:
        info = {"n_gpus": world, "clips_per_step_per_rank": args.clips, "steps": args.steps, "host_cpus": os.cpu_count(),
                "bytes_per_step_per_rank": {"h2d_f32": n_f32, "h2d_i16": n_i16, "d2h": n_out},
                "gpu": torch.cuda.get_device_name(0), "rows": rows, "notes": notes}
        try:
            import subprocess
            info["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-3000:]
            info["lscpu"] = subprocess.run(["bash", "-c", "lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread'"],
                                           capture_output=True, text=True, timeout=20).stdout
        except Exception:  # noqa: BLE001
            pass
        print(json.dumps(info, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
