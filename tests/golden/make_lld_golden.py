"""Regenerates tests/golden/lld_golden_v1.npz: the OpenSMILE-path descriptors and functionals of seeded synthetic
clips as computed by the numpy restatement oracle/lld_oracle.py (PARITY UNPINNED: no SMILExtract binary, no OpenSMILE output in
the reference repository), frozen so that the restatement cannot drift silently and the CUDA path is compared with numbers that
do not depend on the numpy build of the GPU box.  Both descriptor sets: the first slice (MFCC / energy / ZCR, mean + stddev,
56 columns) and the widest one (30 contours x 2 x 12 functionals = 720 columns).  Run from the repo root:
    python tests/golden/make_lld_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lld_oracle as lo  # noqa: E402
from robust_speech_analysis_framework_b200.lld_extractor import functional_names  # noqa: E402
from robust_speech_analysis_framework_b200.synth import synth_clip  # noqa: E402


def main():
    rng = np.random.default_rng(21)
    clips = [synth_clip(640 + i, d).numpy() for i, d in enumerate([2.0, 1.13, 0.6])]
    clips += [np.zeros(4800, np.int16), (rng.normal(scale=2500.0, size=9000)).astype(np.int16), np.zeros(100, np.int16)]
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    fun56, rows56 = lo.extract(pcm, off, 16000.0)
    fun720, rows720 = lo.extract(pcm, off, 16000.0, descriptor_set=1, functional_set=1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lld_golden_v1.npz")
    np.savez_compressed(out, pcm=pcm, offsets=off, functionals_56=fun56, functionals_720=fun720,
                        frames_720=np.concatenate([r for r in rows720 if len(r)]),
                        frame_counts=np.array([len(r) for r in rows720]),
                        names_56=np.array(functional_names()), names_720=np.array(functional_names(descriptor_set=1, functional_set=1)))
    print("wrote", out, fun56.shape, fun720.shape, [len(r) for r in rows720])


if __name__ == "__main__":
    main()
