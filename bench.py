#!/usr/bin/env python
"""Headline benchmark: clips/sec (10 s, 16 kHz) of logmel + CRNN inference (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W]           # B200 arm (this repo's kernels)
  python bench.py --impl reference [--steps K] [--warmup W]     # reference algorithm on the host CPU

A step = one pass of the whole hot path (waveform -> log-mel -> Cnn9 -> bi-GRU -> frame-attention pooling)
over one batch of synthetic clips: BASELINE.json configs[1], Cnn_9layers_Gru_FrameAtt, 16 kHz, batch 1024 per
GPU.  For N > 1 each rank owns its own 1024-clip shard (weak scaling, no collective on the data path) and the
framewise/clipwise outputs are gathered to rank 0 with NCCL inside the timed step.  One JSON line on stdout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_TYPE = "Cnn_9layers_Gru_FrameAtt"
SR, N_FFT, HOP, FMIN, FMAX = 16000, 512, 160, 25, 7000
CLIP_SAMPLES = 160000
CONV_GFLOP_TC = 26.031 - 0.0738  # tensor-core conv layers per clip (SURVEY.md 8d minus conv_block1.conv1)
METRIC = "clips/sec (10 s, 16 kHz) logmel+CRNN inference"


def config_index():
    """BASELINE.json configs[1] = the GRU model (the headline); configs[2] = the Transformer model, 512 clips per GPU."""
    return 1 if MODEL_TYPE == "Cnn_9layers_Gru_FrameAtt" else 2


def load_conv_traffic():
    """DRAM bytes of the seven conv launches of one 148-clip micro-batch, from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f)
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            pk = json.load(f)
        return {"tflops_sustained": float(pk.get("bf16_tflops_sustained", 1410.6)),
                "tflops_burst": float(pk.get("bf16_tflops", 1678.0)), "hbm_gbs": float(pk.get("hbm_gbs", 6537.6)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed regions run.  The sampler is started before the
    warm-up (nvidia-smi needs a few hundred ms to come up); only samples taken between mark_begin() and mark_end()
    -- i.e. under the benchmark's load -- are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []   # (monotonic time, text)
        self.t_begin = None
        self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.monotonic()

    def mark_end(self):
        self.t_end = time.monotonic()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        t1 = (self.t_end if self.t_end is not None else time.monotonic()) + 0.06  # a sample reports the interval before it
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if ts < t0 or ts > t1:
                continue
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": "device-resident + end-to-end timed regions"}


def cpu_reference_throughput(steps, warmup, batch=32, device="cpu"):
    """The reference algorithm (oracle port of pytorch/models.py forward) on the host CPU, all cores.  With
    device="cuda" (`--ref-device cuda`, an extra row, never the reference arm the driver runs) the same float32 torch
    ops run eagerly on the GPU: the incumbent a user of the reference has today."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import sed_oracle
    from sed_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.synthetic_state_dict(MODEL_TYPE, SR)
    wave = synth.synthetic_waveform(batch, CLIP_SAMPLES, seed=1234)
    on_gpu = device != "cpu"
    if on_gpu:
        sd = {k: v.to(device) for k, v in sd.items()}
        wave = wave.to(device)

    def sync():
        if on_gpu:
            torch.cuda.synchronize()

    for _ in range(warmup):
        sed_oracle.model_forward(sd, wave, MODEL_TYPE, N_FFT, HOP)
    sync()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        out = sed_oracle.model_forward(sd, wave, MODEL_TYPE, N_FFT, HOP)
        if on_gpu:
            out["framewise_output"].cpu()
        sync()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    where = ("float32 torch eager ops on %s (cudnn tf32 %s)" % (torch.cuda.get_device_name(0),
                                                                 torch.backends.cudnn.allow_tf32)) if on_gpu else \
        "float32 torch CPU ops, %d threads" % cores
    return {"value": batch * steps / total, "unit": "clips/s", "cores": cores, "kind": "port",
            "sample": "%d steps of batch %d x 10 s clips, %s" % (steps, batch, where),
            "ms_per_step": 1e3 * total / steps, "batch": batch}


def pm_spans(B, micro_batch):
    return list(range(0, B, micro_batch))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    warmup = max(0, args.warmup)
    batch = args.ref_batch
    cb = cpu_reference_throughput(steps, warmup, batch, args.ref_device)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "clips/s", "n_gpus": args.gpus,
        "gpus_used": 0 if args.ref_device == "cpu" else 1, "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s logmel 16k batch %d per GPU (BASELINE.json configs[%d]), "
                               "10 s clips, seeded synthetic checkpoint" % (MODEL_TYPE, args.batch, config_index()),
                   "batch_per_gpu": args.batch,
                   "sample": "each step is %d clips of that workload on the host CPU (same clips, same checkpoint)"
                             % batch},
        "cpu_baseline": {"value": cb["value"], "unit": "clips/s", "cores": cb["cores"], "kind": "port",
                         "sample": cb["sample"]},
        "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from sed_b200 import capi, engine, synth
    from sed_b200 import dist as sdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: everything any library prints to fd 1 (NCCL's version banner, ...) is sent
    # to stderr, the result line is written to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the B200 arm)")
    full_affinity = os.sched_getaffinity(0)
    numa = "off" if os.environ.get("SED_NO_NUMA_BIND") else sdist.bind_host_to_gpu(local_rank)  # before pinned buffers
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL banners / warnings never reach stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    capi.load()

    B = args.batch
    steps, warmup = max(1, args.steps), max(3, args.warmup)
    sd = synth.synthetic_state_dict(MODEL_TYPE, SR)
    pm = engine.PackedModel(sd, MODEL_TYPE, N_FFT, HOP, dev, precision=args.precision)
    wave_host = synth.synthetic_waveform(B, CLIP_SAMPLES, seed=1234, rank=rank).pin_memory()
    wave = wave_host.to(dev)

    def step_device():
        out = pm.forward(wave, micro_batch=args.micro_batch, variant=args.variant)
        if world > 1:
            sdist.gather_outputs(out, dst=0)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        step_device()
    barrier()

    # ---------------- timed region 1: inputs resident in HBM ----------------
    if rank == 0:
        sampler.mark_begin()
    capi.reset_launches()
    pm.conv_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = capi.launches()
    conv_ms = sum(a.elapsed_time(b) for a, b, _ in pm.conv_events)
    conv_clips = sum(n for _, _, n in pm.conv_events)
    pm.conv_events = None
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = t.item()

    # ---------------- timed region 2: end to end with HOST buffers ----------------
    def step_host():
        out = pm.forward_host(wave_host, micro_batch=args.micro_batch, variant=args.variant)
        return out

    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(steps):
        res = step_host()
    h1.record()
    barrier()
    e2e_ms = max(h0.elapsed_time(h1), 0.0)
    wall_ms = 1e3 * (time.perf_counter() - t0)
    t = torch.tensor([max(e2e_ms, wall_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = t.item()
    d2h = res["clipwise_output"].numel() * 4 + res["framewise_output"].numel() * 4

    # ---------------- extra: the same end-to-end call fed with int16 PCM ----------------
    wave_i16 = torch.round(wave_host * 32767.0).to(torch.int16).pin_memory()
    for _ in range(2):
        pm.forward_host(wave_i16, micro_batch=args.micro_batch, variant=args.variant)
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        pm.forward_host(wave_i16, micro_batch=args.micro_batch, variant=args.variant)
    barrier()
    t = torch.tensor([1e3 * (time.perf_counter() - t0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_i16_ms = t.item()
    clocks = None
    if rank == 0:
        sampler.mark_end()
        clocks = sampler.stop()

    if rank == 0:
        peaks = load_peaks()
        traffic = load_conv_traffic()
        value = world * B * steps / (elapsed_ms / 1e3)
        # variant 4 runs conv_block1.conv1 inside the first timed launch: the timed group is the whole conv stack
        conv_gflop = 26.031 if args.variant == 4 else CONV_GFLOP_TC
        conv_tflops = conv_gflop * 1e9 * conv_clips / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
        n_conv_launch = 7 * len(pm_spans(B, args.micro_batch)) * steps
        line = {
            "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": elapsed_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic",
            "config": {"workload": "%s logmel 16k batch %d per GPU (BASELINE.json configs[%d]), "
                                   "10 s clips, seeded synthetic checkpoint" % (MODEL_TYPE, B, config_index()),
                       "batch_per_gpu": B, "micro_batch": args.micro_batch, "parallelism": "dp%d (batch shards, "
                       "NCCL gather of outputs to rank 0)" % world,
                       "l2": "inputs (%.0f MB/step) and activations (>6 GB/micro-batch) exceed the 126 MB L2" %
                             (B * CLIP_SAMPLES * 4 / 1e6)},
            "e2e": {"value": world * B * steps / (e2e_ms / 1e3), "unit": "clips/s",
                    "h2d_bytes_per_step": B * CLIP_SAMPLES * 4, "d2h_bytes_per_step": d2h,
                    "api": "PackedModel.forward_host (pinned host f32 waveform in, host clipwise/framewise out)",
                    "host_affinity": numa,
                    "int16_input_value": world * B * steps / (e2e_i16_ms / 1e3),
                    "int16_h2d_bytes_per_step": B * CLIP_SAMPLES * 2},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peaks["tflops_sustained"],
                         "unit": "TFLOP/s", "frac": conv_tflops / peaks["tflops_sustained"],
                         "traffic": traffic["dram_bytes_total"] * B / traffic["clips"] if traffic else None,
                         "traffic_note": "DRAM bytes of the 7 conv launches of one step (%d clips): dram__bytes_read+write "
                                         "from ncu --set full of a 148-clip launch group (profiles/r01_ncu_full_conv_umma2"
                                         "_raw.csv, r01_ncu_full_conv_block1_tc_raw.csv: %.3e B), scaled by clips; "
                                         "algorithmic activation bytes (each layer input read once + output written "
                                         "once) for the same launches: %.3e" % (B, traffic["dram_bytes_total"] if traffic
                                                                                else 0.0, B * 21888256.0),
                         "kernel": "conv_block1_tc_kernel + 6 x conv_umma2_kernel (7 tcgen05 cta_group::2 implicit-GEMM "
                                   "launches per micro-batch, %.3f GFLOP/clip algorithmic)" % conv_gflop,
                         "peak_source": peaks["source"] + ", sustained bf16; burst %.1f" % peaks["tflops_burst"],
                         "launches_timed": n_conv_launch, "conv_ms_per_step": conv_ms / steps},
        }
        if world == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, full_affinity)  # the CPU baseline gets every host core again
            cb = cpu_reference_throughput(steps=3, warmup=1, batch=args.ref_batch)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="clips per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=1036, help="cap of a conv-stack launch group (engine default)")
    ap.add_argument("--variant", type=int, default=4)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--model-type", default=MODEL_TYPE,
                    choices=["Cnn_9layers_Gru_FrameAtt", "Cnn_9layers_Transformer_FrameAtt"],
                    help="default = the headline config; the Transformer model is BASELINE config 3 (use --batch 512)")
    ap.add_argument("--ref-batch", type=int, default=32)
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="cpu = the reference arm (default); cuda = extra row: the same torch ops eagerly on the GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    globals()["MODEL_TYPE"] = args.model_type
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
