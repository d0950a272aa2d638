"""One 60 s clip (BASELINE.json configs[0]) through the device-resident entry, for ncu captures of the latency-bound kernels.
python tools/profile_single.py [seconds] [passes]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robust_speech_analysis_framework_b200 import _lib
from robust_speech_analysis_framework_b200.synth import synth_batch

dur = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pcm, off = synth_batch(1, dur, "cuda", start_index=5000)
off = off.numpy().astype(np.int64)
out = torch.empty((1, 25), dtype=torch.float64, device="cuda")
st = torch.empty(1, dtype=torch.int32, device="cuda")
ex = _lib.Extractor(0)
for _ in range(passes):
    ex.extract_device(pcm.data_ptr(), off, out.data_ptr(), st.data_ptr())
torch.cuda.synchronize()
ex.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ex.extract_device(pcm.data_ptr(), off, out.data_ptr(), st.data_ptr())
e1.record()
torch.cuda.synchronize()
print(f"one {dur:g} s clip: {e0.elapsed_time(e1):.2f} ms")
for name, (ms, spans) in sorted(ex.profile_report().items(), key=lambda kv: -kv[1][0])[:8]:
    print(f"  {name:40s} {ms:9.3f} ms")
