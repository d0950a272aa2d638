"""One extraction pass bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off (run on the GPU box).

python tools/profile_pass.py [n_clips] [seconds] [warmup_passes]
ncu --profile-from-start off --metrics <...> --csv --log-file out.csv python tools/profile_pass.py 32 30
Only the profiled pass is captured: every kernel of the 25-column pipeline exactly once per chunk.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robust_speech_analysis_framework_b200 import _lib
from robust_speech_analysis_framework_b200.synth import synth_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dur = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pcm, off = synth_batch(n, dur, "cuda", start_index=0, unique=min(n, 32))
off = off.numpy().astype(np.int64)
out = torch.empty((n, 25), dtype=torch.float64, device="cuda")
st = torch.empty(n, dtype=torch.int32, device="cuda")
ex = _lib.Extractor(0)
for _ in range(warm):
    ex.extract_device(pcm.data_ptr(), off, out.data_ptr(), st.data_ptr())
torch.cuda.synchronize()
torch.cuda.profiler.start()
ex.extract_device(pcm.data_ptr(), off, out.data_ptr(), st.data_ptr())
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"profiled pass: {n} clips x {dur:g} s = {n * dur:g} audio-s, launches so far {ex.launch_count}, "
      f"nan columns {int(torch.isnan(out).any(dim=0).sum())}")
