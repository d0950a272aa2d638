"""Which recordings of the ragged set (BASELINE.json configs[2]) have NaN columns, with their status words, batch vs alone
(run on the GPU box).  Finding of round 2: ONE recording (index 170, 76 s) has F0 / slope / tilt / CPP / moments NaN with status
0x60 in the batch and alone -- its main pitch pass finds no voiced frame -- and the CPU port returns the same row
(`python tools/dbg_cfg2.py`; the CPU side was checked with oracle.mshds_oracle.extract on the same samples)."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
from robust_speech_analysis_framework_b200 import _lib
from robust_speech_analysis_framework_b200.synth import synth_batch
FS=16000
durs = bench.ragged_durations(230)
uniq = 48
base, boff = synth_batch(uniq, 600.0, "cuda", start_index=7000)
base = base.cpu().numpy(); boff = boff.numpy()
clips = [base[boff[i % uniq]: boff[i % uniq] + int(round(durs[i] * FS))] for i in range(len(durs))]
off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
pcm = np.concatenate(clips)
ex = _lib.Extractor(0)
out, st = ex.extract_host(pcm, off)
bad = np.where(np.isnan(out).any(axis=1))[0]
print("clips with NaN:", len(bad), bad[:20])
for i in bad[:10]:
    print(i, "dur", durs[i], "status", hex(st[i]), "nan cols", np.where(np.isnan(out[i]))[0])
    alone, st1 = ex.extract_host(clips[i], np.array([0, len(clips[i])], np.int64))
    print("   alone: status", hex(st1[0]), "nan cols", np.where(np.isnan(alone[0]))[0])
