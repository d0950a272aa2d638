#!/usr/bin/env python
"""bench.py -- MSHDS extraction throughput (audio-seconds per second) on N B200s of one node.

Default workload (BASELINE.json configs[1]): 1,000 x 30 s synthetic 16 kHz voiced clips per GPU, all 25 MSHDS columns.
A "step" is one pass of the whole hot path over that batch.

  value     whole-job audio-s/s with the int16 batch already resident in HBM (device pointers through the C ABI),
            timed with CUDA events on the (non-default) torch stream the library is told to run on, max over ranks.
  e2e       same metric through the host-buffer C-ABI call a user of the drop-in makes (pinned host int16 in, host
            float64 out): H2D of the batch and D2H of the feature matrix inside the timed region.
  parity    rows of the TIMED batch are compared with the CPU oracle run on the same int16 samples (the first
            --check-clips clips): `parity_rows_checked`, `max_rel_diff`, `decision_flips`; a mismatch fails the run.
  roofline  dominant kernel FUNCTION (per-stage CUDA-event spans inside the library, summed over the passes that run the
            same kernel): algorithmic bytes (32,000 B per audio-second of int16 + 200 B per clip) / its time
            against the measured HBM peak -- tiny by construction for a ~1e4 FLOP/byte float64 pipeline -- and, as the
            binding figure, float64 FLOP/s against the DFMA peak measured in this run (`mshds_fp64_peak`); FLOP counts per
            audio-second come from ncu SASS op counters (tools/flop_model.py -> profiles/r02_flop_model.json).
  cpu_baseline  the reference's CPU path on the host cores: the real extractor when praat-parselmouth is importable
            (oracle/reference_probe.py, kind "reference"), else the CPU oracle port (kind "port"); bounded sample.

Other BASELINE.json configs (`--config 0|2|3|4`, one JSON line each; measurement cases, not the driver's default):
  0  one 60 s clip: latency of one recording, GPU and CPU
  2  ~230 ragged recordings U(60, 600) s, STRONG scaling: LPT-sharded over the ranks through sharding.extract_sharded with
     the real Extractor, per-rank busy time -> load imbalance
  3  100,000 x 2 s short clips (launch / packing stress), strong scaling over the ranks
  4  STFT / MFCC parameter sweep of the OpenSMILE-LLD slice (n_fft 512/1024/2048, 40/80 mel bands)

`--impl reference` times the CPU path alone with the same metric / config (rank 0 only under torchrun).
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
FLOP_MODEL = os.path.join(ROOT, "profiles", "r02_flop_model.json")

# tolerance groups of tests/test_gpu_parity.py (rtol, atol) per column
REL_TOL = {"speechrate": (1e-12, 1e-12), "pitch": (1e-7, 1e-7), "continuous": (1e-6, 1e-9), "formant": (1e-5, 1e-6)}
GROUP_OF = ["speechrate"] * 5 + ["pitch"] * 2 + ["continuous"] * 6 + ["formant"] * 8 + ["continuous"] * 4

# stage function (span name up to '[') -> substrings of the kernel names that run inside it (for the FLOP model)
STAGE_KERNELS = {
    "pitch_ac_frames": ["k_ac_frames", "k_pitch_frames<0", "k_pitch_frames<false", "k_pitch_frames<(bool)0"],
    "pitch_cc_frames": ["k_pitch_frames<1", "k_pitch_frames<true", "k_pitch_frames<(bool)1"],
    "k_pitch_refine": ["k_pitch_refine"],
    "k_hnr_refine": ["k_hnr_refine"],
    "cepstrogram_frames": ["k_cepstrogram"],
    "cpps_frames": ["k_cpp_frames"],
    "formant_burg_frames": ["k_formant_frames"],
    "spectrogram_moments": ["k_spec_"],
    "resample_clip_10k": ["k_fft_", "k_sinc_"],
    "resample_segments_10k": ["k_fft_", "k_sinc_"],
    "ltas": ["k_ltas"],
}


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


# ------------------------------------------------------------------------------------------------ CPU side
def cpu_reference_run(pcm: np.ndarray, off: np.ndarray, threads: int):
    """The reference's CPU path on given int16 clips -> (features [n,25], seconds, kind).  The real extractor
    (unmodified /root/reference/src/mshds_extractor.py on praat-parselmouth) when importable, else the CPU oracle port."""
    from oracle import reference_probe as probe
    if probe.available():
        t0 = time.perf_counter()
        feats, _ = probe.extract(pcm, off, FS, processes=threads)
        return feats, time.perf_counter() - t0, "reference"
    from oracle import mshds_oracle as orc
    orc.lib()
    t0 = time.perf_counter()
    feats, _ = orc.extract(pcm, off, float(FS), nthreads=threads)
    return feats, time.perf_counter() - t0, "port"


def compare_rows(got: np.ndarray, want: np.ndarray):
    """Per-column-group comparison of feature rows -> dict(parity_ok, max_rel_diff, decision_flips, ...)."""
    nan_mismatch = int((np.isnan(got) != np.isnan(want)).sum())
    both = ~np.isnan(got) & ~np.isnan(want)
    rel = np.zeros_like(got)
    rel[both] = np.abs(got[both] - want[both]) / np.maximum(np.abs(want[both]), 1e-12)
    bad_cols = []
    for k in range(25):
        rtol, atol = REL_TOL[GROUP_OF[k]]
        b = both[:, k]
        if not np.all(np.abs(got[b, k] - want[b, k]) <= atol + rtol * np.abs(want[b, k])):
            bad_cols.append(k)
    sr_diff = (~np.isclose(got[:, :5], want[:, :5], rtol=1e-12, atol=1e-12, equal_nan=True)).any(axis=1)
    return {
        "parity_rows_checked": int(got.shape[0]),
        "max_rel_diff": float(rel.max()) if rel.size else 0.0,
        "max_rel_diff_column": int(np.unravel_index(rel.argmax(), rel.shape)[1]) if rel.size else None,
        "decision_flips": int(sr_diff.sum()),       # clips whose count-derived speech-rate columns differ (a flipped decision)
        "nan_mismatches": nan_mismatch,
        "columns_out_of_tolerance": bad_cols,
        "parity_ok": bool(not bad_cols and nan_mismatch == 0),
        "tolerances": "tests/test_gpu_parity.py REL_TOL groups (speech-rate 1e-12, pitch 1e-7, continuous 1e-6, formant 1e-5)",
    }


def workload_name(args):
    return {0: "1 x 60 s synthetic 16 kHz voiced clip (BASELINE.json configs[0])",
            1: f"{args.clips} x {args.seconds:g} s synthetic 16 kHz voiced clips per GPU, all 25 MSHDS columns "
               f"({'BASELINE.json configs[1]' if (args.clips == 1000 and args.seconds == 30.0) else 'non-default shape'})",
            2: f"{args.ragged_clips} ragged synthetic recordings U(60, 600) s, sharded over the ranks (BASELINE.json configs[2])",
            3: f"{args.short_clips} x 2 s short clips sharded over the ranks (BASELINE.json configs[3])",
            4: "OpenSMILE-LLD slice, n_fft 512/1024/2048 x 40/80 mel bands (BASELINE.json configs[4])"}[args.config]


def run_reference(args):
    """--impl reference: the reference's CPU path on host cores, bounded sample per step, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from robust_speech_analysis_framework_b200.synth import synth_batch
    cores = os.cpu_count() or 1
    if args.config == 0:
        n_clips, seconds = 1, 60.0
    elif args.config == 3:
        n_clips, seconds = max(cores, 64), 2.0
    else:
        n_clips, seconds = max(1, min(cores, args.ref_clips if args.ref_clips > 0 else cores)), args.seconds
    times, kind = [], "port"
    for step in range(args.warmup + args.steps):
        pcm, off = synth_batch(n_clips, seconds, "cpu", start_index=900000 + 100 * step)
        _, dt, kind = cpu_reference_run(pcm.numpy(), off.numpy(), cores)
        if step >= args.warmup:
            times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = n_clips * seconds / (ms / 1000.0)
    sample = f"{n_clips} clips x {seconds:g} s per step (a bounded sample of the workload), {cores} host threads"
    from oracle import reference_probe as probe
    line = {
        "impl": "reference", "metric": "mshds_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if args.config in (2, 3) else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "note": ("reference = unmodified src/mshds_extractor.py on praat-parselmouth" if kind == "reference" else
                            "reference = CPU oracle port of src/mshds_extractor.py + Praat (" + probe.why_not() + ")")},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU side helpers
class Ctx:
    """torch / distributed / extractor set-up shared by the configs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from robust_speech_analysis_framework_b200 import _lib
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ex = _lib.Extractor(self.local_rank)
        # a dedicated non-default stream: the library issues on it, the events that time it are recorded on it
        self.stream = torch.cuda.Stream(device=self.dev)
        self.ex.set_stream(self.stream.cuda_stream)
        if args.chunk_log2:
            self.ex.set_chunk_samples(1 << args.chunk_log2)

    def timed(self, fn, steps):
        torch, dist = self.torch, self.dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for _ in range(steps):
                fn()
            e1.record(self.stream)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        own = ms
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, own

    def gather_scalars(self, v: float):
        if self.world == 1:
            return [v]
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def by_function(stages):
    """Sum the main-stream spans that run the same kernel function ('name[...]' -> 'name'); '~' (side stream) and indented
    (sub-span) entries are not costs of the main stream."""
    agg = {}
    for name, (ms, cnt) in stages.items():
        if name.startswith("~") or name.startswith(" "):
            continue
        fn = name.split("[")[0]
        a = agg.setdefault(fn, [0.0, 0])
        a[0] += ms; a[1] += cnt
    return agg


def load_flop_model():
    try:
        return json.load(open(FLOP_MODEL))
    except Exception:
        return None


def flops_of(model, substrings):
    if not model:
        return None
    tot = 0.0
    for k, v in model.get("kernels", {}).items():
        if any(s in k for s in substrings):
            tot += v["flop"]
    return tot / model["audio_seconds"] if tot > 0 else None


# ------------------------------------------------------------------------------------------------ config 1 (default)
def bench_default(args):
    C = Ctx(args)
    torch, dist, ex = C.torch, C.dist, C.ex
    from robust_speech_analysis_framework_b200.synth import synth_batch
    import ctypes as CT
    rank, world, dev = C.rank, C.world, C.dev

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
    lib, h = ex._lib, ex._h

    def step_device():
        ex.extract_device(pcm_d.data_ptr(), off_np, out_d.data_ptr(), st_d.data_ptr())
        if world > 1:
            dist.all_gather(gathered, out_d)          # only the small feature matrix crosses NVLink

    def step_host():
        rc = lib.mshds_extract(h, CT.c_void_p(pcm_h.data_ptr()), off_np.ctypes.data, n, FS, CT.c_void_p(out_h.data_ptr()),
                               CT.c_void_p(st_h.data_ptr()), 0)
        if rc != 0:
            raise RuntimeError(lib.mshds_last_error(h).decode())
        if world > 1:
            out_d.copy_(out_h, non_blocking=True)
            dist.all_gather(gathered, out_d)

    warm = max(args.warmup, 3)
    with torch.cuda.stream(C.stream):
        for _ in range(warm):
            step_device()
    sampler = ClockSampler(C.local_rank)
    if rank == 0:
        sampler.start()
    ex.profile(True)
    l0 = ex.launch_count
    ms_total, _ = C.timed(step_device, args.steps)
    launches = ex.launch_count - l0
    stages = ex.profile_report()
    ex.profile(False)
    feats_dev = out_d.cpu().numpy().copy()           # rows produced by the timed device-resident steps
    with torch.cuda.stream(C.stream):
        step_host()
    ms_e2e, _ = C.timed(step_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    feats_host = out_h.numpy().copy()

    ms_step = ms_total / args.steps
    value = world * audio_s / (ms_step / 1000.0)
    e2e_value = world * audio_s / (ms_e2e / args.steps / 1000.0)
    nan_cols = int(np.isnan(feats_dev).any(axis=0).sum())
    if rank != 0:
        C.close()
        return 0

    fp64_peak = ex.fp64_peak_tflops()
    hbm_gbs, peak_src = measured_peaks()
    fn = by_function(stages)
    dom = max(fn, key=lambda k: fn[k][0]) if fn else None
    dom_ms, dom_launches = fn[dom] if dom else (float("nan"), 0)
    alg_bytes_step = audio_s * BYTES_PER_AUDIO_SECOND + n * BYTES_PER_CLIP_OUT
    passes_per_step = dom_launches / args.steps if dom else None           # launches of that function per step (passes x chunks)
    # one launch of a frame kernel covers one chunk of the batch: its algorithmic bytes are the chunk's share
    chunks_per_step = max(1, int(np.ceil(off_np[-1] / float(1 << (args.chunk_log2 or 27)))))
    alg_bytes_launch = alg_bytes_step / chunks_per_step
    avg_launch_ms = dom_ms / dom_launches if dom_launches else None
    achieved = alg_bytes_launch / (avg_launch_ms / 1000.0) / 1e9 if dom_launches else None
    model = load_flop_model()
    flop_step = model["flop_per_audio_second"] * audio_s if model else None
    dom_flop_as = flops_of(model, STAGE_KERNELS.get(dom, [dom or "?"]))
    traffic = None
    if model and dom:
        tr = [v.get("dram_bytes_per_launch") for k, v in model.get("kernels", {}).items()
              if any(s in k for s in STAGE_KERNELS.get(dom, [dom])) and v.get("dram_bytes_per_launch")]
        traffic = float(np.mean(tr)) if tr else None
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
        "frac": (achieved / hbm_gbs) if achieved else None, "traffic": traffic, "peak_source": peak_src,
        "kernel_ms_per_step": dom_ms / args.steps if dom else None,
        "kernel_share_of_step": (dom_ms / args.steps) / ms_step if dom else None,
        "launches_per_step": passes_per_step, "avg_launch_ms": avg_launch_ms, "algorithmic_bytes_per_launch": alg_bytes_launch,
        "fp64": {
            "peak_tflops": fp64_peak, "peak_source": "DFMA issue peak measured in this run (mshds_fp64_peak)",
            "flop_per_audio_second": model["flop_per_audio_second"] if model else None,
            "flop_source": (f"ncu SASS op counters (2*dfma + dmul + dadd), {model.get('workload')}, profiles/r02_flop_model.json"
                            if model else "no FLOP model committed"),
            "pipeline_tflops": flop_step / (ms_step / 1000.0) / 1e12 if flop_step else None,
            "pipeline_frac": flop_step / (ms_step / 1000.0) / 1e12 / fp64_peak if flop_step and fp64_peak else None,
            "kernel_tflops": (dom_flop_as * audio_s / (dom_ms / args.steps / 1000.0) / 1e12) if dom_flop_as and dom else None,
            "kernel_frac": (dom_flop_as * audio_s / (dom_ms / args.steps / 1000.0) / 1e12 / fp64_peak) if dom_flop_as and dom and fp64_peak else None,
        },
        "note": "float64 compute-bound pipeline (~1e4 FLOP per compulsory byte): the HBM fraction is tiny by construction; "
                "the binding roofline is `fp64` (measured FLOP rate / measured DFMA peak)",
    }
    stage_table = {k: round(v[0] / args.steps, 3) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])}
    fn_table = {k: round(v[0] / args.steps, 3) for k, v in sorted(fn.items(), key=lambda kv: -kv[1][0])}
    if args.stages:
        for k, v in stage_table.items():
            print(f"{v:10.3f} ms/step  {k}", file=sys.stderr)

    # ---- parity of the timed rows + CPU baseline: the reference's CPU path on the first clips of THIS batch
    cpu, parity = None, None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ncheck = min(n, args.check_clips if args.check_clips > 0 else max(32, cores))
        sub_pcm = pcm_h.numpy()[: off_np[ncheck]]
        want, dt, kind = cpu_reference_run(sub_pcm, off_np[: ncheck + 1], cores)
        cpu = {"value": float(off_np[ncheck]) / FS / dt, "unit": "audio-s/s", "cores": cores, "kind": kind,
               "sample": f"the first {ncheck} of the {n} timed clips ({off_np[ncheck] / FS:g} audio-s), {cores} host threads, {dt:.1f} s"}
        parity = compare_rows(feats_dev[:ncheck], want)
        ph = compare_rows(feats_host[:ncheck], want)
        parity["e2e_rows_ok"] = ph["parity_ok"]
        parity["device_and_host_legs_identical"] = bool(np.array_equal(feats_dev, feats_host, equal_nan=True))
        parity["checked_against"] = "real reference (praat-parselmouth)" if kind == "reference" else "CPU oracle port (parity unpinned)"

    line = {
        "metric": "mshds_audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "audio_seconds_per_gpu": audio_s,
                   "l2": f"int16 batch {pcm_d.numel() * 2 / 1e9:.2f} GB per GPU > 126 MB L2 (no flush needed)",
                   "unique_clips": args.unique or n, "synth_seconds": round(gen_s, 1), "nan_columns": nan_cols},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(pcm_d.numel() * 2 + off_np.nbytes),
                "d2h_bytes_per_step": int(n * (25 * 8 + 4)), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "stages_ms_per_step": stage_table,
        "functions_ms_per_step": fn_table,
        "stages_note": "CUDA-event spans inside the library; '~' = issued on the side stream underneath main-stream kernels "
                       "(wall time while sharing the SMs, not additive)",
    }
    print(json.dumps(line))
    C.close()
    if parity is not None and not (parity["parity_ok"] and parity["e2e_rows_ok"]):
        print("bench.py: PARITY FAILURE -- the timed rows differ from the CPU reference path", file=sys.stderr)
        return 1
    return 0


# ------------------------------------------------------------------------------------------------ config 0: one 60 s clip
def bench_single_clip(args):
    C = Ctx(args)
    torch, ex = C.torch, C.ex
    from robust_speech_analysis_framework_b200.synth import synth_batch
    pcm_d, off = synth_batch(1, 60.0, C.dev, start_index=5000)
    off_np = off.numpy().astype(np.int64)
    pcm_h = torch.empty(pcm_d.shape, dtype=torch.int16, pin_memory=True)
    pcm_h.copy_(pcm_d)
    out_d = torch.empty((1, 25), dtype=torch.float64, device=C.dev)
    st_d = torch.empty(1, dtype=torch.int32, device=C.dev)
    host = pcm_h.numpy()

    def step_device():
        ex.extract_device(pcm_d.data_ptr(), off_np, out_d.data_ptr(), st_d.data_ptr())

    res = {}

    def step_host():
        res["f"], _ = ex.extract_host(host, off_np)

    warm = max(args.warmup, 3)
    with torch.cuda.stream(C.stream):
        for _ in range(warm):
            step_device()
    ex.profile(True)
    l0 = ex.launch_count
    ms, _ = C.timed(step_device, args.steps)
    launches = ex.launch_count - l0
    stages = ex.profile_report()
    ex.profile(False)
    ms_h, _ = C.timed(step_host, args.steps)
    if C.rank != 0:
        C.close()
        return 0
    ms_step, ms_host = ms / args.steps, ms_h / args.steps
    cores = os.cpu_count() or 1
    want, dt, kind = cpu_reference_run(host, off_np, 1)
    parity = compare_rows(res["f"], want)
    line = {
        "metric": "mshds_audio_seconds_per_second", "value": 60.0 / (ms_step / 1e3), "unit": "audio-s/s", "n_gpus": 1,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": {"workload": workload_name(args), "latency_ms_device_resident": ms_step,
                                                         "latency_ms_host_buffers": ms_host},
        "e2e": {"value": 60.0 / (ms_host / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": int(host.nbytes + 16), "d2h_bytes_per_step": 204,
                "ms_per_step": ms_host},
        "gpu_launches": int(launches),
        "cpu_baseline": {"value": 60.0 / dt, "unit": "audio-s/s", "cores": 1, "kind": kind,
                         "sample": f"the same 60 s clip, 1 host thread (the reference is a serial loop), {dt:.2f} s; host has {cores} cores"},
        "parity": parity,
        "stages_ms_per_step": {k: round(v[0] / args.steps, 3) for k, v in sorted(stages.items(), key=lambda kv: -kv[1][0])},
    }
    print(json.dumps(line))
    C.close()
    return 0 if parity["parity_ok"] else 1


# ------------------------------------------------------------------------------------------------ configs 2 / 3: strong scaling
def ragged_durations(n, seed=20251018):
    rng = np.random.default_rng(seed)
    return np.round(rng.uniform(60.0, 600.0, n), 3)


def bench_sharded(args):
    """Strong scaling of a fixed clip set over the ranks through sharding.extract_sharded with the real Extractor (host
    buffers in, gathered matrix out on rank 0).  Reports per-rank busy time: the limiter is load imbalance + serial tails."""
    C = Ctx(args)
    torch, ex = C.torch, C.ex
    from robust_speech_analysis_framework_b200 import sharding
    from robust_speech_analysis_framework_b200.synth import synth_batch
    if args.config == 2:
        durs = ragged_durations(args.ragged_clips)
        uniq = min(args.ragged_clips, args.unique or 48)
        # distinct recordings are expensive to synthesise at 5 min each: `unique` distinct signals, every clip its own length
        base, boff = synth_batch(uniq, 600.0, C.dev, start_index=7000)
        base = base.cpu().numpy(); boff = boff.numpy()
        clips = [base[boff[i % uniq]: boff[i % uniq] + int(round(durs[i] * FS))] for i in range(len(durs))]
    else:
        n = args.short_clips
        uniq = min(n, args.unique or 2000)
        base, boff = synth_batch(uniq, 2.0, C.dev, start_index=9000)
        base = base.cpu().numpy(); boff = boff.numpy()
        clips = [base[boff[i % uniq]: boff[i % uniq + 1]] for i in range(n)]
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    pcm = np.concatenate(clips)
    del clips
    audio_s = float(off[-1]) / FS
    n = len(off) - 1
    busy = {}

    def compute(sub_pcm, sub_off):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(C.stream)
        r = ex.extract_host(sub_pcm, sub_off)
        e1.record(C.stream)
        e1.synchronize()
        busy["ms"] = busy.get("ms", 0.0) + e0.elapsed_time(e1)
        return r

    out = {}

    def step():
        out["f"], out["s"] = sharding.extract_sharded(pcm, off, compute, C.rank, C.world)

    warm = max(args.warmup, 1)
    with torch.cuda.stream(C.stream):
        for _ in range(warm):
            step()
    busy.clear()
    l0 = ex.launch_count
    t0 = time.perf_counter()
    ms, own = C.timed(step, args.steps)
    wall = time.perf_counter() - t0
    launches = ex.launch_count - l0
    per_rank_busy = C.gather_scalars(busy.get("ms", 0.0) / args.steps)
    parts = sharding.lpt_assign(np.diff(off), C.world)
    per_rank_audio = [float(sum(off[i + 1] - off[i] for i in p)) / FS for p in parts]
    if C.rank != 0:
        C.close()
        return 0
    ms_step = ms / args.steps
    feats = out["f"]
    line = {
        "metric": "mshds_audio_seconds_per_second", "value": audio_s / (ms_step / 1e3), "unit": "audio-s/s", "n_gpus": C.world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "clips": n, "audio_seconds_total": audio_s, "unique_signals": uniq,
                   "timed_region": "sharding.extract_sharded: LPT split, host packing of the rank's clips, H2D, extraction, all_gather of the "
                                   "[n_local, 26] matrix, scatter to input order on rank 0",
                   "nan_columns": int(np.isnan(feats).any(axis=0).sum())},
        "e2e": {"value": audio_s / (ms_step / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": int(off[-1] * 2), "d2h_bytes_per_step": int(n * 204),
                "ms_per_step": ms_step},
        "gpu_launches": int(launches),
        "per_rank": {"audio_seconds": per_rank_audio, "gpu_busy_ms_per_step": per_rank_busy,
                     "imbalance_max_over_mean": float(max(per_rank_busy) / max(np.mean(per_rank_busy), 1e-9)),
                     "host_overhead_ms_per_step": ms_step - max(per_rank_busy)},
        "wall_s": wall,
    }
    print(json.dumps(line))
    C.close()
    return 0


# ------------------------------------------------------------------------------------------------ config 4: LLD sweep
def bench_lld_sweep(args):
    C = Ctx(args)
    torch, ex = C.torch, C.ex
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200.synth import synth_batch
    clips, seconds = args.lld_clips, args.lld_seconds
    pcm_d, off = synth_batch(clips, seconds, C.dev, start_index=C.rank * clips, unique=min(clips, 16))
    off_np = off.numpy().astype(np.int64)
    audio_s = float(off_np[-1]) / FS
    out_d = torch.empty((clips, 56), dtype=torch.float64, device=C.dev)
    pcm_h = pcm_d.cpu().numpy()
    hbm, src = measured_peaks()
    rows = []
    for n_fft in (512, 1024, 2048):
        for n_mel in (40, 80):
            params = dict(n_fft=n_fft, n_mel=n_mel)

            def dev_step():
                ex.lld_extract_device(pcm_d.data_ptr(), off_np, out_d.data_ptr(), 16000, **params)

            with torch.cuda.stream(C.stream):
                for _ in range(3):
                    dev_step()
            ms, _ = C.timed(dev_step, args.steps)
            ms /= args.steps
            ms_h, _ = C.timed(lambda: ex.lld_extract(pcm_h, off_np, 16000, **params), max(1, args.steps // 2))
            ms_h /= max(1, args.steps // 2)
            alg = audio_s * 32000 + clips * 56 * 8
            rows.append({"n_fft": n_fft, "n_mel": n_mel, "audio_s_per_s": C.world * audio_s / (ms / 1e3),
                         "e2e_audio_s_per_s": C.world * audio_s / (ms_h / 1e3), "ms_per_step": ms,
                         "hbm_gbs": alg / (ms / 1e3) / 1e9, "hbm_frac": alg / (ms / 1e3) / 1e9 / hbm})
    # the widest descriptor / functional set built so far (Androids.conf defaults: n_fft 512, 26 mel bands): 720 columns
    full = dict(descriptor_set=1, functional_set=1)
    out_full = torch.empty((clips, 720), dtype=torch.float64, device=C.dev)

    def full_step():
        ex.lld_extract_device(pcm_d.data_ptr(), off_np, out_full.data_ptr(), 16000, **full)

    with torch.cuda.stream(C.stream):
        for _ in range(3):
            full_step()
    ms_full, _ = C.timed(full_step, args.steps)
    ms_full /= args.steps
    full_row = {"columns": 720, "audio_s_per_s": C.world * audio_s / (ms_full / 1e3), "ms_per_step": ms_full}
    if C.rank != 0:
        C.close()
        return 0
    t0 = time.perf_counter()
    nref = min(clips, 2)
    fun_ref, _ = lo.extract(pcm_h[: off_np[nref]], off_np[: nref + 1], 16000.0, n_fft=1024, n_mel=40)
    cpu_s = time.perf_counter() - t0
    got, _, _ = ex.lld_extract(pcm_h[: off_np[nref]], off_np[: nref + 1], 16000, n_fft=1024, n_mel=40)
    best = max(rows, key=lambda r: r["audio_s_per_s"])
    line = {
        "metric": "lld_audio_seconds_per_second", "value": best["audio_s_per_s"], "unit": "audio-s/s", "n_gpus": C.world, "steps": args.steps,
        "warmup": 3, "ms_per_step": best["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "clips_per_gpu": clips, "seconds_per_clip": seconds,
                   "columns": "MFCC 1-12 + RMS energy + ZCR, smoothed + deltas, mean / stddev (56 columns; parity unpinned: no SMILExtract binary)"},
        "e2e": {"value": best["e2e_audio_s_per_s"], "unit": "audio-s/s", "h2d_bytes_per_step": int(pcm_h.nbytes), "d2h_bytes_per_step": int(clips * 56 * 8)},
        "sweep": rows,
        "androids_720_columns": full_row,
        "roofline": {"bound": "hbm", "achieved": best["hbm_gbs"], "peak": hbm, "unit": "GB/s", "frac": best["hbm_frac"], "peak_source": src,
                     "traffic": None},
        "cpu_baseline": {"value": nref * seconds / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port",
                         "sample": f"{nref} x {seconds:g} s, numpy restatement (oracle/lld_oracle.py), n_fft 1024 / 40 bands"},
        "parity": {"max_abs_diff_vs_numpy_oracle": float(np.nanmax(np.abs(got - fun_ref)))},
    }
    print(json.dumps(line))
    C.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[0, 1, 2, 3, 4], help="BASELINE.json configs index (default 1 = the headline)")
    ap.add_argument("--clips", type=int, default=1000, help="clips per GPU (config 1)")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--unique", type=int, default=0, help="synthesise only this many distinct clips (0 = all distinct / config default)")
    ap.add_argument("--ref-clips", type=int, default=0, help="clips per step of --impl reference (0 = one per core)")
    ap.add_argument("--check-clips", type=int, default=0, help="rows of the timed batch checked against the CPU path (0 = max(32, cores))")
    ap.add_argument("--ragged-clips", type=int, default=230)
    ap.add_argument("--short-clips", type=int, default=100000)
    ap.add_argument("--lld-clips", type=int, default=256)
    ap.add_argument("--lld-seconds", type=float, default=600.0)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (and with it the parity check); profiling runs only")
    ap.add_argument("--stages", action="store_true", help="print the per-stage timing table to stderr")
    ap.add_argument("--chunk-log2", type=int, default=0, help="override the library's chunk size (log2 of samples per chunk)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return 0
    fn = {0: bench_single_clip, 1: bench_default, 2: bench_sharded, 3: bench_sharded, 4: bench_lld_sweep}[args.config]
    return fn(args)


if __name__ == "__main__":
    sys.exit(main())
