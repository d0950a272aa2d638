"""Frame-level descriptors + functionals of the OpenSMILE path, first slice.

The reference's second handcrafted extractor, /root/reference/src/opensmile_extractor.py:9-103, shells out to the external
SMILExtract binary once per file with Androids.conf and returns one row of 911 functionals per recording.  This module
keeps that function's shape -- DataFrame of file paths in, one row per recording out, 'filename' first, NaN row + printed
message for a file that cannot be processed (:89-99) -- for the part of the component graph built so far: MFCC 1-12,
RMS energy and zero-crossing rate per 25 ms / 10 ms frame (Androids.conf:73-132), cIntensity intensity / loudness (:134-140)
and 14 of the 16 cSpectral descriptors (:257-282), smoothed (cContourSmoother) and with regression deltas (cDeltaRegression)
like the 'lld' / 'lld_de' levels of the config, with the twelve functionals of functL1 (:349-366) over the recording: 720 of the
911 columns.  Not built: SHS pitch + Viterbi smoother, jitter / shimmer, psySharpness, spectralHarmonicity.  Column names follow OpenSMILE's '<lld>_<functional>' pattern.  The arithmetic runs in
libmshds_b200.so (mshds_lld_extract); there is no CPU fallback.
"""
from __future__ import annotations

import os

import numpy as np

from . import mshds_extractor as _mx


SPECTRAL_NAMES = ["pcm_intensity", "pcm_loudness", "pcm_fftMag_fband250-650", "pcm_fftMag_fband1000-4000",
                  "pcm_fftMag_spectralRollOff25.0", "pcm_fftMag_spectralRollOff50.0", "pcm_fftMag_spectralRollOff75.0",
                  "pcm_fftMag_spectralRollOff90.0", "pcm_fftMag_spectralFlux", "pcm_fftMag_spectralCentroid",
                  "pcm_fftMag_spectralEntropy", "pcm_fftMag_spectralVariance", "pcm_fftMag_spectralSkewness",
                  "pcm_fftMag_spectralKurtosis", "pcm_fftMag_spectralSlope", "pcm_fftMag_spectralFlatness"]
FUNCTIONALS_2 = ["amean", "stddev"]
FUNCTIONALS_12 = ["max", "min", "range", "maxPos", "minPos", "amean", "linregc1", "linregc2", "linregerrQ", "stddev", "skewness",
                  "kurtosis"]                             # Androids.conf functL1 (:349-366): Extremes, Regression, Moments


def lld_names(n_mfcc: int = 12, smooth_win: int = 3, delta_win: int = 2, descriptor_set: int = 0):
    """OpenSMILE-style contour names: 'mfcc_sma[1]', 'pcm_RMSenergy_sma', ..., then the '_de' regression deltas."""
    sma = "_sma" if smooth_win > 1 else ""
    extra = SPECTRAL_NAMES if descriptor_set else []
    base = [f"mfcc{sma}[{i}]" for i in range(1, n_mfcc + 1)] + [f"pcm_RMSenergy{sma}", f"pcm_zcr{sma}"] + [f"{e}{sma}" for e in extra]
    if delta_win > 0:
        base = base + [f"mfcc{sma}_de[{i}]" for i in range(1, n_mfcc + 1)] + [f"pcm_RMSenergy{sma}_de", f"pcm_zcr{sma}_de"] + \
            [f"{e}{sma}_de" for e in extra]
    return base


def functional_names(n_mfcc: int = 12, smooth_win: int = 3, delta_win: int = 2, descriptor_set: int = 0, functional_set: int = 0):
    names = lld_names(n_mfcc, smooth_win, delta_win, descriptor_set)
    return [f"{n}_{fn}" for fn in (FUNCTIONALS_12 if functional_set else FUNCTIONALS_2) for n in names]


def extract_lld_functionals(input_df, audio_file_column='filepath', verbose=True, device: int = 0, max_batch_seconds: float = 7200.0,
                            **params):
    """One row per recording: 'filename' + the functionals of every descriptor contour.  `params` override mshds_lld_params
    fields (frame_size, frame_step, preemph, n_fft, n_mel, mel_lo, mel_hi, n_mfcc, cep_lifter, smooth_win, delta_win,
    descriptor_set, functional_set).  Defaults here are the widest set built so far -- descriptor_set = 1 (MFCC 1-12, RMS
    energy, ZCR, intensity, loudness, 14 cSpectral descriptors) and functional_set = 1 (the twelve functionals of Androids.conf
    functL1): 30 contours x 2 (deltas) x 12 = 720 of the 911 columns of Androids.conf; pass descriptor_set=0, functional_set=0
    for the first slice (56 columns).  Recordings are sent to the device in
    batches of at most `max_batch_seconds` of audio per sampling rate (the library allocates for a whole call)."""
    import pandas as pd

    params.setdefault("descriptor_set", 1)
    params.setdefault("functional_set", 1)
    ex = _mx.get_extractor(device)
    paths = [row[audio_file_column] for _, row in input_df.iterrows()]
    filenames = [os.path.basename(p) for p in paths]
    cols = functional_names(int(params.get("n_mfcc", 12)), int(params.get("smooth_win", 3)), int(params.get("delta_win", 2)),
                            int(params["descriptor_set"]), int(params["functional_set"]))
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
