"""Development aid (run on the GPU box, usually under ncu): one extract call over N synthetic clips of D seconds.

python tools/profile_small.py [n_clips] [seconds] [calls] [sample_rate]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robust_speech_analysis_framework_b200 import _lib
from robust_speech_analysis_framework_b200.synth import synth_clip

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dur = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 1
fs = int(sys.argv[4]) if len(sys.argv) > 4 else 16000
uniq = min(n, 32)
base = [synth_clip(1000 + i, dur, device="cuda", fs=fs).cpu().numpy() for i in range(uniq)]
clips = [base[i % uniq] for i in range(n)]
pcm = np.concatenate(clips)
off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
ex = _lib.Extractor(0)
for k in range(calls):
    ex.profile(True)
    t = time.time()
    out, st = ex.extract_host(pcm, off, fs)
    dt = time.time() - t
    print(f"call {k}: {dt:.3f} s  {n * dur / dt:.0f} audio-s/s  launches {ex.launch_count}  nan cols {int(np.isnan(out).any(axis=0).sum())}")
    if k == calls - 1:
        for name, (ms, spans) in ex.profile_report().items():
            print(f"  {name:45s} {ms:10.3f} ms  {spans}")
