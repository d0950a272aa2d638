"""Frame-level descriptors + functionals of the OpenSMILE path, first slice.

The reference's second handcrafted extractor, /root/reference/src/opensmile_extractor.py:9-103, shells out to the external
SMILExtract binary once per file with Androids.conf and returns one row of 911 functionals per recording.  This module
keeps that function's shape -- DataFrame of file paths in, one row per recording out, 'filename' first, NaN row + printed
message for a file that cannot be processed (:89-99) -- for the part of the component graph built so far: MFCC 1-12,
RMS energy and zero-crossing rate per 25 ms / 10 ms frame (Androids.conf:73-132), smoothed (cContourSmoother) and with
regression deltas (cDeltaRegression) like the 'lld' / 'lld_de' levels of the config, with their mean and standard
deviation over the recording.  Column names follow OpenSMILE's '<lld>_<functional>' pattern.  The arithmetic runs in
libmshds_b200.so (mshds_lld_extract); there is no CPU fallback.
"""
from __future__ import annotations

import os

import numpy as np

from . import mshds_extractor as _mx


def lld_names(n_mfcc: int = 12, smooth_win: int = 3, delta_win: int = 2):
    """OpenSMILE-style contour names: 'mfcc_sma[1]', 'pcm_RMSenergy_sma', ..., then the '_de' regression deltas."""
    sma = "_sma" if smooth_win > 1 else ""
    base = [f"mfcc{sma}[{i}]" for i in range(1, n_mfcc + 1)] + [f"pcm_RMSenergy{sma}", f"pcm_zcr{sma}"]
    if delta_win > 0:
        base = base + [f"mfcc{sma}_de[{i}]" for i in range(1, n_mfcc + 1)] + [f"pcm_RMSenergy{sma}_de", f"pcm_zcr{sma}_de"]
    return base


def functional_names(n_mfcc: int = 12, smooth_win: int = 3, delta_win: int = 2):
    names = lld_names(n_mfcc, smooth_win, delta_win)
    return [f"{n}_amean" for n in names] + [f"{n}_stddev" for n in names]


def extract_lld_functionals(input_df, audio_file_column='filepath', verbose=True, device: int = 0, max_batch_seconds: float = 7200.0,
                            **params):
    """One row per recording: 'filename' + mean / stddev of every descriptor.  `params` override mshds_lld_params fields
    (frame_size, frame_step, preemph, n_fft, n_mel, mel_lo, mel_hi, n_mfcc, cep_lifter).  Recordings are sent to the device in
    batches of at most `max_batch_seconds` of audio per sampling rate (the library allocates for a whole call)."""
    import pandas as pd

    ex = _mx.get_extractor(device)
    paths = [row[audio_file_column] for _, row in input_df.iterrows()]
    filenames = [os.path.basename(p) for p in paths]
    cols = functional_names(int(params.get("n_mfcc", 12)), int(params.get("smooth_win", 3)), int(params.get("delta_win", 2)))
    feats = np.full((len(paths), len(cols)), np.nan)
    by_rate = {}
    for i, path in enumerate(paths):
        try:
            pcm, fs = _mx.read_wav_mono_int16(path)
            if len(pcm) == 0:
                raise _mx.AudioLoadError("empty sound")
            by_rate.setdefault(fs, ([], []))
            by_rate[fs][0].append(i)
            by_rate[fs][1].append(pcm)
        except Exception as e:
            if verbose:
                print(f"  - ERROR processing {filenames[i]}: {e}")
    def run(idx, clips, fs):
        offs = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
        try:
            fun, _, _ = ex.lld_extract(np.concatenate(clips), offs, fs, **params)
            feats[np.asarray(idx)] = fun
        except Exception as e:
            if len(idx) > 1:              # isolate the failure: one recording per call
                for i, c in zip(idx, clips):
                    run([i], [c], fs)
                return
            import warnings
            warnings.warn(f"LLD extraction failed for '{filenames[idx[0]]}': {e}", RuntimeWarning)
            if verbose:
                print(f"  - ERROR processing {filenames[idx[0]]}: {e}")

    for fs, (idx, clips) in by_rate.items():
        start, acc = 0, 0
        for k, c in enumerate(clips):
            acc += len(c)
            if acc >= max_batch_seconds * fs or k == len(clips) - 1:
                run(idx[start:k + 1], clips[start:k + 1], fs)
                start, acc = k + 1, 0
    df = pd.DataFrame(feats, columns=cols)
    df.insert(0, "filename", filenames)
    return df
