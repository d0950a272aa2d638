#!/usr/bin/env python
"""bench.py -- MSHDS extraction throughput (audio-seconds per second) on N B200s of one node.

Workload (BASELINE.json configs[1]): 1,000 x 30 s synthetic 16 kHz voiced clips per GPU, all 25 MSHDS columns.
A "step" is one pass of the whole hot path over that batch.

  value     whole-job audio-s/s with the int16 batch already resident in HBM (device pointers through the C ABI),
            timed with CUDA events on the stream the kernels run on, max over ranks.
  e2e       same metric through the host-buffer C-ABI call a user of the drop-in makes (pinned host int16 in, host
            float64 out): H2D of the batch and D2H of the feature matrix inside the timed region.
  roofline  dominant main-stream stage (per-stage CUDA-event timing inside the library): algorithmic bytes (32,000 B per audio-second
            of int16 + 200 B per clip) / stage time, against the measured HBM peak.  The workload is
            ~1e4 FLOP per byte, so this fraction is tiny by construction; the fp64 figure next to it is the binding one.
  cpu_baseline  the CPU oracle (a restatement of the reference's Praat calls, kind "port") on a bounded sample, all cores.

`--impl reference` times that CPU path alone (the reference itself cannot run: praat-parselmouth is not installable
offline) with the same metric / config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 16000
BYTES_PER_AUDIO_SECOND = 2 * FS      # int16 ingest
BYTES_PER_CLIP_OUT = 25 * 8


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_throughput(n_clips: int, seconds: float, threads: int, start_index: int = 900000):
    """audio-s/s of the CPU oracle on n_clips x seconds synthetic clips with `threads` OpenMP threads."""
    from oracle import mshds_oracle as orc
    from robust_speech_analysis_framework_b200.synth import synth_batch
    pcm, off = synth_batch(n_clips, seconds, "cpu", start_index=start_index)
    pcm, off = pcm.numpy(), off.numpy()
    orc.lib()
    t0 = time.perf_counter()
    orc.extract(pcm, off, float(FS), nthreads=threads)
    dt = time.perf_counter() - t0
    return n_clips * seconds / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the real one needs praat-parselmouth) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_clips = max(1, min(cores, args.ref_clips if args.ref_clips > 0 else cores))
    times = []
    for step in range(args.warmup + args.steps):
        v, dt = cpu_oracle_throughput(n_clips, args.seconds, cores, start_index=900000 + 100 * step)
        if step >= args.warmup:
            times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = n_clips * args.seconds / (ms / 1000.0)
    sample = f"{n_clips} clips x {args.seconds:g} s per step (of the {args.clips} x {args.seconds:g} s workload), {cores} OpenMP threads"
    line = {
        "impl": "reference", "metric": "mshds_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.clips} x {args.seconds:g} s synthetic 16 kHz voiced clips, 25 MSHDS columns (BASELINE.json configs[1])",
                   "note": "reference = CPU oracle port of src/mshds_extractor.py + Praat (praat-parselmouth not installable offline)"},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=1000, help="clips per GPU")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--unique", type=int, default=0, help="synthesise only this many distinct clips (0 = all distinct)")
    ap.add_argument("--ref-clips", type=int, default=0, help="clips per step of the CPU baseline sample (0 = one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stages", action="store_true", help="print the per-stage timing table to stderr")
    ap.add_argument("--chunk-log2", type=int, default=0, help="override the library's chunk size (log2 of samples per chunk)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from robust_speech_analysis_framework_b200 import _lib
    from robust_speech_analysis_framework_b200.synth import synth_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # ---- synthetic batch, generated on the device, int16, packed; a pinned host copy for the e2e leg
    t0 = time.perf_counter()
    pcm_d, off = synth_batch(args.clips, args.seconds, dev, start_index=rank * args.clips, unique=args.unique or None)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    off_np = off.numpy().astype(np.int64)
    n = args.clips
    audio_s = float(off_np[-1]) / FS
    pcm_h = torch.empty(pcm_d.shape, dtype=torch.int16, pin_memory=True)
    pcm_h.copy_(pcm_d)
    out_d = torch.empty((n, 25), dtype=torch.float64, device=dev)
    st_d = torch.empty(n, dtype=torch.int32, device=dev)
    out_h = torch.empty((n, 25), dtype=torch.float64, pin_memory=True)
    st_h = torch.empty(n, dtype=torch.int32, pin_memory=True)
    gathered = [torch.empty_like(out_d) for _ in range(world)] if world > 1 else None

    ex = _lib.Extractor(local_rank)
    stream = torch.cuda.current_stream()
    ex.set_stream(stream.cuda_stream)
    if args.chunk_log2:
        ex.set_chunk_samples(1 << args.chunk_log2)
    lib, h = ex._lib, ex._h
    import ctypes as C

    def step_device():
        ex.extract_device(pcm_d.data_ptr(), off_np, out_d.data_ptr(), st_d.data_ptr())
        if world > 1:
            dist.all_gather(gathered, out_d)          # only the small feature matrix crosses NVLink

    def step_host():
        rc = lib.mshds_extract(h, C.c_void_p(pcm_h.data_ptr()), off_np.ctypes.data, n, FS, C.c_void_p(out_h.data_ptr()),
                               C.c_void_p(st_h.data_ptr()), 0)
        if rc != 0:
            raise RuntimeError(lib.mshds_last_error(h).decode())
        if world > 1:
            out_d.copy_(out_h, non_blocking=True)
            dist.all_gather(gathered, out_d)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident leg
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ex.profile(True)
    l0 = ex.launch_count
    ms_total = timed(step_device, args.steps)
    launches = ex.launch_count - l0
    stages = ex.profile_report()
    ex.profile(False)
    # ---- end-to-end leg (host buffers through the C ABI)
    step_host()
    ms_e2e = timed(step_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    ms_step = ms_total / args.steps
    value = world * audio_s / (ms_step / 1000.0)
    e2e_value = world * audio_s / (ms_e2e / args.steps / 1000.0)

    # sanity: the result must be a full, finite feature matrix (no skipped work)
    feats = out_d.cpu().numpy()
    nan_cols = int(np.isnan(feats).any(axis=0).sum())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_gbs, peak_src = measured_peaks()
    dom, dom_ms = None, 0.0
    for name, (ms, cnt) in stages.items():
        # "~" spans are side-stream work (Viterbi, pulse walks): they overlap the main stream, their wall time is not a cost;
        # indented spans are parts of another span
        if name.startswith("~") or name.startswith(" "):
            continue
        if ms > dom_ms:
            dom, dom_ms = name, ms
    alg_bytes_step = audio_s * BYTES_PER_AUDIO_SECOND + n * BYTES_PER_CLIP_OUT
    dom_ms_step = dom_ms / args.steps if dom else float("nan")
    dom_launches = stages[dom][1] if dom else 0                      # one launch per chunk of <= 2^27 samples
    # algorithmic bytes of ONE launch = the chunk's share of the batch; achieved = that / the average launch duration
    alg_bytes_launch = alg_bytes_step * args.steps / dom_launches if dom_launches else None
    avg_launch_ms = dom_ms / dom_launches if dom_launches else None
    achieved = alg_bytes_launch / (avg_launch_ms / 1000.0) / 1e9 if dom_launches else None
    traffic, fp64_pct = None, None
    try:   # dram__bytes_read+write and fp64 pipe utilisation of the dominant kernel from the committed ncu capture
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_dominant_kernel.json")))
        if prof.get("kernel") == dom:
            traffic = prof.get("dram_bytes_per_launch")
            fp64_pct = prof.get("fp64_pipe_pct_of_peak")
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
        "frac": (achieved / hbm_gbs) if achieved else None, "traffic": traffic,
        "peak_source": peak_src, "kernel_ms_per_step": dom_ms_step, "kernel_share_of_step": dom_ms_step / ms_step if dom else None,
        "launches_per_step": dom_launches / args.steps if dom else None, "avg_launch_ms": avg_launch_ms,
        "algorithmic_bytes_per_launch": alg_bytes_launch, "fp64_pipe_pct_of_peak": fp64_pct,
        "note": "float64 compute-bound pipeline (~1e4 FLOP per compulsory byte): the HBM fraction is tiny by construction; "
                "the binding figure is the FP64 pipe utilisation (DESIGN.md, profiles/)",
    }
    stage_table = {k: round(v[0] / args.steps, 3) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])}
    if args.stages:
        for k, v in stage_table.items():
            print(f"{v:10.3f} ms/step  {k}", file=sys.stderr)

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nref = args.ref_clips if args.ref_clips > 0 else cores
        v, dt = cpu_oracle_throughput(nref, args.seconds, cores)
        cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
               "sample": f"{nref} of the {n} clips' kind ({nref} x {args.seconds:g} s), CPU oracle with {cores} OpenMP threads, {dt:.1f} s"}

    line = {
        "metric": "mshds_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{n} x {args.seconds:g} s synthetic 16 kHz voiced clips per GPU, all 25 MSHDS columns ({'BASELINE.json configs[1]' if (n == 1000 and args.seconds == 30.0) else 'non-default shape'})",
                   "audio_seconds_per_gpu": audio_s, "l2": f"int16 batch {pcm_d.numel() * 2 / 1e9:.2f} GB per GPU > 126 MB L2 (no flush needed)",
                   "unique_clips": args.unique or n, "synth_seconds": round(gen_s, 1), "nan_columns": nan_cols},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(pcm_d.numel() * 2 + off_np.nbytes),
                "d2h_bytes_per_step": int(n * (25 * 8 + 4)), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stages_ms_per_step": stage_table,
        "stages_note": "CUDA-event spans inside the library; '~' = issued on the side stream underneath main-stream kernels "
                       "(wall time while sharing the SMs, not additive)",
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
