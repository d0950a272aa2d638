"""Development aid (run on the GPU box): stage-by-stage comparison of libmshds_b200 against the CPU oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mshds_oracle as orc
from robust_speech_analysis_framework_b200 import _lib
from robust_speech_analysis_framework_b200.synth import synth_clip

durs = [float(a) for a in sys.argv[1:]] or [4.0, 5.00006, 3.0, 6.5]
clips = [synth_clip(i, d).numpy() for i, d in enumerate(durs)]
pcm = np.concatenate(clips)
off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
ex = _lib.Extractor(0)
t = time.time()
out, st = ex.extract_host(pcm, off)
print("gpu extract s", time.time() - t, "status", st, "launches", ex.launch_count)
t = time.time()
ref, rst = orc.extract(pcm, off, 16000.0, nthreads=8)
print("oracle s", time.time() - t, "status", rst)
np.set_printoptions(linewidth=200, precision=9)
for k, name in enumerate(_lib.FEATURE_NAMES):
    d = np.abs(out[:, k] - ref[:, k]) / np.maximum(np.abs(ref[:, k]), 1e-300)
    print(f"{name:26s} gpu {out[:, k]} ref {ref[:, k]} rel {np.nanmax(d) if not np.all(np.isnan(d)) else np.nan:.3e}")


def cmp(name, got, want, tol=1e-9):
    n = min(len(got), len(want))
    if len(got) != len(want):
        print(f"  {name}: LENGTH {len(got)} vs {len(want)}")
    if n == 0:
        return
    g, w = got[:n], want[:n]
    both_nan = np.isnan(g) & np.isnan(w)
    d = np.where(both_nan, 0.0, np.abs(g - w))
    d = np.where(np.isnan(d), np.inf, d)
    bad = np.nonzero(d > tol * np.maximum(1.0, np.abs(w)))[0]
    print(f"  {name}: n={n} maxabs={np.max(d):.3e} nbad={len(bad)} first_bad={bad[:5]}")
    for b in bad[:3]:
        print(f"     [{b}] gpu={g[b]!r} ref={w[b]!r}")


for c in range(len(durs)):
    x = orc.pcm_to_float(clips[c])
    cls = ex.debug_fetch("class", c, np.int32)
    fl, ce, fb = orc.pitch_values(x)
    print(f"clip {c}: n={len(x)} class gpu={cls} oracle floor={fl} ceil={ce} fallback={fb}")
    pw = orc.pitch(x, 16000.0, 0, 0.005, 50.0, 3.0, 15, 0.03, 0.45, 0.01, 0.35, 0.14, 600.0)
    cmp("pitch_wide_f", ex.debug_fetch("pitch_wide_f", c), pw["freq"])
    pm = orc.pitch(x, 16000.0, 0, 0.005, fl, 3.0, 15, 0.03, 0.45, 0.01, 0.35, 0.14, ce)
    cmp("pitch_main_f", ex.debug_fetch("pitch_main_f", c), pm["freq"])
    cmp("pitch_main_s", ex.debug_fetch("pitch_main_s", c), pm["strength"])
    ic, _ = orc.intensity(x, 16000.0, fl, 0.005)
    cmp("intensity_main", ex.debug_fetch("intensity_main", c), ic)
    ph = orc.pitch(x, 16000.0, 2, 0.005, fl, 4.5, 15, 0.1, 0.0, 0.0, 0.0, 0.0, 8000.0)
    hr = ex.debug_fetch("hnr_r", c)
    want = np.where(ph["freq"] == 0, np.nan, ph["strength"])
    cmp("hnr_r", hr, want)
    psr = orc.pitch(x, 16000.0, 0, 0.02, 30.0, 3.0, 4, 0.03, 0.25, 0.01, 0.35, 0.25, 450.0)
    cmp("pitch_sr_f", ex.debug_fetch("pitch_sr_f", c), psr["freq"])
    isr, _ = orc.intensity(x, 16000.0, 50.0, 0.016)
    cmp("intensity_sr", ex.debug_fetch("intensity_sr", c), isr)
    plt_ = orc.pitch(x, 16000.0, 0, 0.0, fl, 3.0, 15, 0.03, 0.45, 0.01, 0.35, 0.14, ce)
    cmp("pitch_ltas_f", ex.debug_fetch("pitch_ltas_f", c), plt_["freq"])
    pul = orc.pulses(x, 16000.0, 0, 0.0, fl, 3.0, 0.45, ce)
    cmp("pulses_ltas", ex.debug_fetch("pulses_ltas", c), pul, tol=1e-9)
    lb = orc.ltas(x, 16000.0, fl, ce)
    if lb is not None:
        cmp("ltas_bands", ex.debug_fetch("ltas_bands", c), lb[0], tol=1e-9)
    rs, rx1 = orc.resample(x, 16000.0, 10000.0, 500)
    cmp("resampled10k", ex.debug_fetch("resampled10k", c), rs, tol=1e-9)
    fo = orc.formants(x, 16000.0)
    gn = ex.debug_fetch("formant_n", c, np.int32)
    print("  formant_n equal:", np.array_equal(gn, fo["n"]), len(gn), len(fo["n"]))
    gf = ex.debug_fetch("formant_f", c).reshape(-1, 5)
    gb = ex.debug_fetch("formant_b", c).reshape(-1, 5)
    cmp("formant_f", gf.ravel(), fo["f"].ravel(), tol=1e-7)
    cmp("formant_b", gb.ravel(), fo["bw"].ravel(), tol=1e-6)
    pcc = orc.pitch(x, 16000.0, 2, 0.005, fl, 1.0, 15, 0.03, 0.45, 0.01, 0.35, 0.14, ce)
    cmp("pitch_cc_f", ex.debug_fetch("pitch_cc_f", c), pcc["freq"])
    cmp("pulses_fmt", ex.debug_fetch("pulses_fmt", c), orc.pulses(x, 16000.0, 2, 0.005, fl, 1.0, 0.45, ce))
    cmp("pulses_cpp", ex.debug_fetch("pulses_cpp", c), orc.pulses(x, 16000.0, 0, 0.005, fl, 3.0, 0.3, ce))
