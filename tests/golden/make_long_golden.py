"""Generates tests/golden/long_clips_golden_v1.npz: CPU-oracle features of an Androids-scale ragged batch (BASELINE.json
configs[2]: 1-10 min recordings).  The oracle needs minutes for these, so the GPU suite compares against the stored values
instead of running it on the GPU box.  The clips are re-synthesised from their seeds; a checksum guards the comparison.

    python tests/golden/make_long_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SPECS = [(900, 600.0), (901, 61.30006), (902, 200.5)]        # (synthesis index, seconds)


def make_batch():
    from robust_speech_analysis_framework_b200.synth import synth_clip
    clips = [synth_clip(i, d).numpy() for i, d in SPECS]
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    return pcm, off


if __name__ == "__main__":
    from oracle import mshds_oracle as orc
    pcm, off = make_batch()
    feats, status = orc.extract(pcm, off, 16000.0, nthreads=len(SPECS))
    np.savez_compressed(os.path.join(HERE, "long_clips_golden_v1.npz"), features=feats, status=status, offsets=off,
                        sha256=hashlib.sha256(pcm.tobytes()).hexdigest())
    print(feats[:, :8], status)
