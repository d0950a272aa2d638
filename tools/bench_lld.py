"""Throughput of the OpenSMILE LLD slice (mshds_lld_extract) on one B200 -- a side measurement, not the bench contract.

python tools/bench_lld.py [--clips 1000] [--seconds 30] [--n-fft 0] [--n-mel 26] [--steps 5]
Prints one JSON line: audio-s/s device-resident and through host buffers, achieved algorithmic GB/s against the measured
HBM peak, and the numpy restatement timed on one host core for a bounded sample.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=1000)
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--n-fft", type=int, default=0)
    ap.add_argument("--n-mel", type=int, default=26)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    import torch
    from bench import measured_peaks
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200 import _lib
    from robust_speech_analysis_framework_b200.synth import synth_batch
    dev = torch.device("cuda", 0)
    pcm_d, off = synth_batch(a.clips, a.seconds, dev, unique=min(a.clips, 64))
    off_np = off.numpy().astype(np.int64)
    audio_s = float(off_np[-1]) / 16000.0
    ex = _lib.Extractor(0)
    stream = torch.cuda.current_stream()
    ex.set_stream(stream.cuda_stream)
    out_d = torch.empty((a.clips, 56), dtype=torch.float64, device=dev)
    params = dict(n_fft=a.n_fft, n_mel=a.n_mel)
    pcm_h = pcm_d.cpu().numpy()

    def dev_step():
        ex.lld_extract_device(pcm_d.data_ptr(), off_np, out_d.data_ptr(), 16000, **params)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for _ in range(3):
        dev_step()
    ms = timed(dev_step, a.steps)
    ex.lld_extract(pcm_h, off_np, 16000, **params)
    ms_h = timed(lambda: ex.lld_extract(pcm_h, off_np, 16000, **params), a.steps)
    t0 = time.perf_counter()
    nref = min(a.clips, 4)
    lo.extract(pcm_h[: off_np[nref]], off_np[: nref + 1], 16000.0, **params)
    cpu_s = time.perf_counter() - t0
    hbm, src = measured_peaks()
    alg = audio_s * 32000 + a.clips * 56 * 8
    print(json.dumps({
        "metric": "lld_audio_seconds_per_second", "value": audio_s / (ms / 1e3), "e2e": audio_s / (ms_h / 1e3), "unit": "audio-s/s",
        "ms_per_step": ms, "config": {"workload": f"{a.clips} x {a.seconds:g} s, MFCC 1-12 + energy + zcr, smoothed + deltas, mean/stddev",
                                      "n_fft": a.n_fft, "n_mel": a.n_mel},
        "roofline": {"bound": "hbm", "achieved": alg / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": alg / (ms / 1e3) / 1e9 / hbm, "peak_source": src},
        "cpu_baseline": {"value": nref * a.seconds / cpu_s, "unit": "audio-s/s", "cores": 1, "kind": "port",
                         "sample": f"{nref} x {a.seconds:g} s, numpy restatement"}}))


if __name__ == "__main__":
    main()
