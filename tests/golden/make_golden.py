"""Regenerates tests/golden/mshds_golden_v1.npz.

The reference cannot run anywhere we can reach (praat-parselmouth is absent), so the golden vectors are
outputs of the CPU oracle (oracle/, a restatement of Praat: PARITY UNPINNED) on seeded synthetic clips, frozen here so
that (a) the oracle cannot drift silently and (b) the CUDA path is compared with numbers that do not depend on building
the oracle on the GPU box.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mshds_oracle as orc  # noqa: E402
from robust_speech_analysis_framework_b200.synth import synth_clip  # noqa: E402

DURS = [3.0, 2.50006, 4.2, 3.3]     # second clip has an odd number of samples (frame centres on sample boundaries)


def main():
    clips = [synth_clip(100 + i, d).numpy() for i, d in enumerate(DURS)]
    # edge cases: digital silence, a clip shorter than every analysis window, unvoiced noise
    rng = np.random.default_rng(7)
    clips.append(np.zeros(16000, np.int16))
    clips.append(clips[0][:1200].copy())
    clips.append((rng.normal(scale=800.0, size=24000)).astype(np.int16))
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    feats, status = orc.extract(pcm, off, 16000.0, nthreads=os.cpu_count() or 1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mshds_golden_v1.npz")
    np.savez_compressed(out, pcm=pcm, offsets=off, features=feats, status=status, feature_names=np.array(orc.FEATURE_NAMES))
    print("wrote", out, feats.shape, "status", status)
    np.set_printoptions(linewidth=200, precision=6, suppress=True)
    print(feats)


if __name__ == "__main__":
    main()
