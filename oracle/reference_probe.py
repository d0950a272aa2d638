"""oracle/reference_probe.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Run-time probe for the REAL reference: the unmodified `extract_mshds_features` of
/root/reference/src/mshds_extractor.py:379 on top of praat-parselmouth.

Neither this container nor the GPU box has parselmouth (not in /opt/wheelhouse, no network), so today `find()` returns
None with the reason and every consumer falls back to the CPU port (oracle/*.c, "parity unpinned").  The day a box has the
wheel -- site-packages or `baseline/_ref/` -- the same call sites switch by themselves:

  * tests/test_cpu_host.py::test_oracle_matches_the_real_reference_when_it_is_importable  pins the port against Praat itself,
  * bench.py --impl reference and bench.py's cpu_baseline time the real extractor (`cpu_baseline.kind = "reference"`),
  * tests/golden/make_golden.py --from-reference regenerates the goldens from it.

The reference source is looked for in `baseline/_ref/` (a copy that travels to the GPU box) and, in the build container
only, in /root/reference; it is imported as is with importlib, never copied or edited.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import shutil
import sys
import tempfile
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_DIRS = [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]
_cached = None


def _source_candidates():
    for d in _REF_DIRS:
        for rel in ("src/mshds_extractor.py", "mshds_extractor.py"):
            p = os.path.join(d, rel)
            if os.path.exists(p):
                yield p


def find():
    """(module, None) when the real reference can run here, else (None, reason)."""
    global _cached
    if _cached is not None:
        return _cached
    ref_site = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref_site) and ref_site not in sys.path:
        sys.path.append(ref_site)           # a driver- or user-provided offline install of praat-parselmouth
    try:
        importlib.import_module("parselmouth")
    except Exception as e:                   # ModuleNotFoundError here; an ABI mismatch would also land here
        _cached = (None, f"praat-parselmouth is not importable ({type(e).__name__}: {e}); CPU port used instead")
        return _cached
    src = next(_source_candidates(), None)
    if src is None:
        _cached = (None, "parselmouth imports but src/mshds_extractor.py of the reference was not found under baseline/_ref or /root/reference")
        return _cached
    try:
        spec = importlib.util.spec_from_file_location("reference_mshds_extractor", src)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception as e:
        _cached = (None, f"reference source {src} failed to import: {type(e).__name__}: {e}")
        return _cached
    mod.__reference_source__ = src
    _cached = (mod, None)
    return _cached


def available() -> bool:
    return find()[0] is not None


def why_not() -> str:
    return find()[1] or ""


def write_wavs(pcm: np.ndarray, offsets: np.ndarray, fs: int, directory: str):
    """One 16-bit mono WAV per clip (what the reference API reads, mshds_extractor.py:415); returns the paths."""
    paths = []
    for i in range(len(offsets) - 1):
        p = os.path.join(directory, f"clip_{i:06d}.wav")
        with wave.open(p, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(int(fs))
            w.writeframes(np.ascontiguousarray(pcm[offsets[i]:offsets[i + 1]], dtype="<i2").tobytes())
        paths.append(p)
    return paths


def extract(pcm: np.ndarray, offsets: np.ndarray, fs: int = 16000, processes: int = 1):
    """The real extract_mshds_features on WAVs written to tmpfs -> features [n, 25] in the reference's column order.
    processes > 1 splits the rows over a multiprocessing pool (the reference itself is a serial loop, :408)."""
    mod, reason = find()
    if mod is None:
        raise RuntimeError(reason)
    import pandas as pd
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="mshds_ref_", dir=base)
    try:
        paths = write_wavs(pcm, offsets, fs, d)
        if processes <= 1 or len(paths) < 2:
            df = mod.extract_mshds_features(pd.DataFrame({"filepath": paths}), verbose=False)
        else:
            import multiprocessing as mp
            parts = [paths[k::processes] for k in range(processes) if paths[k::processes]]
            with mp.get_context("fork").Pool(len(parts)) as pool:
                frames = pool.map(_run_part, parts)
            df = pd.concat(frames).set_index("filename").loc[[os.path.basename(p) for p in paths]].reset_index()
        cols = [c for c in df.columns if c != "filename"]
        return df[cols].to_numpy(dtype=np.float64), cols
    finally:
        shutil.rmtree(d, ignore_errors=True)


def _run_part(paths):
    import pandas as pd
    mod, _ = find()
    return mod.extract_mshds_features(pd.DataFrame({"filepath": paths}), verbose=False)
