"""Stage spans of the ragged long-recording set (BASELINE.json configs[2]) on one GPU, for a chunk size given as log2(samples).

python tools/profile_ragged.py [chunk_log2 ...]      (run on the GPU box)
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from robust_speech_analysis_framework_b200 import _lib
from robust_speech_analysis_framework_b200.synth import synth_batch

FS = 16000
durs = bench.ragged_durations(230)
uniq = 48
base, boff = synth_batch(uniq, 600.0, "cuda", start_index=7000)
base = base.cpu().numpy(); boff = boff.numpy()
clips = [base[boff[i % uniq]: boff[i % uniq] + int(round(durs[i] * FS))] for i in range(len(durs))]
off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
pcm = np.concatenate(clips)
audio_s = off[-1] / FS
ex = _lib.Extractor(0)
for lg in [int(a) for a in sys.argv[1:]] or [27]:
    ex.set_chunk_samples(1 << lg)
    ex.extract_host(pcm, off)
    ex.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, st = ex.extract_host(pcm, off)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    rep = ex.profile_report()
    ex.profile(False)
    top = sorted(((k, v[0]) for k, v in rep.items()), key=lambda kv: -kv[1])[:14]
    print(f"chunk 2^{lg}: {audio_s / dt:.0f} audio-s/s ({dt * 1e3:.0f} ms)  " + ", ".join(f"{k} {v:.0f}" for k, v in top), flush=True)
