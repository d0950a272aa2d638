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


# ------------------------------------------------------------------------------------------------ reference-shaped entry point
_CONF_KEYS = {                      # component type -> {config key: (mshds_lld_params field, converter)}
    "cFramer": {"frameSize": ("frame_size", float), "frameStep": ("frame_step", float)},
    "cVectorPreemphasis": {"k": ("preemph", float)},
    "cMelspec": {"lofreq": ("mel_lo", float), "hifreq": ("mel_hi", float), "nBands": ("n_mel", int)},
    "cMfcc": {"lastMfcc": ("n_mfcc", int), "cepLifter": ("cep_lifter", float)},
    "cContourSmoother": {"smaWin": ("smooth_win", int)},
    "cDeltaRegression": {"deltawin": ("delta_win", int)},
}
_BUILT = {"cWaveSource", "cFramer", "cVectorPreemphasis", "cWindower", "cTransformFFT", "cFFTmagphase", "cMelspec", "cMfcc", "cEnergy",
          "cMZcr", "cIntensity", "cSpectral", "cContourSmoother", "cDeltaRegression", "cFunctionals", "cCsvSink", "cComponentManager"}


def parse_smile_config(path: str):
    """The parameters of an OpenSMILE configuration file (Androids.conf) that map onto mshds_lld_params, and the component
    types of the file that this library does not build.  Sections are '[instance:cType]', settings 'key = value', comments
    start with ';' or '//' (/root/reference/Androids.conf:60-420)."""
    params, missing, ctype = {}, [], None
    with open(path, "r", errors="replace") as f:
        for raw in f:
            line = raw.split("//")[0].split(";")[0].strip()
            if not line:
                continue
            if line.startswith("[") and line.endswith("]") and ":" in line:
                ctype = line[1:-1].split(":", 1)[1].strip()
                if ctype not in _BUILT and ctype not in missing:
                    missing.append(ctype)
                continue
            if "=" in line and ctype in _CONF_KEYS:
                key, val = (t.strip() for t in line.split("=", 1))
                if key in _CONF_KEYS[ctype]:
                    field, conv = _CONF_KEYS[ctype][key]
                    try:
                        params[field] = conv(float(val)) if conv is int else conv(val)
                    except ValueError:
                        pass
    return params, missing


def extract_opensmile_features(input_df, opensmile_exe_path=None, config_file_path=None, audio_file_column='filepath', verbose=True,
                               device: int = 0):
    """Drop-in for /root/reference/src/opensmile_extractor.py:9 `extract_opensmile_features`: same arguments and the same
    shape of result -- one row per recording that could be processed (failed files are left out, :89-96), the feature columns
    first and 'filename' last (:86), an empty DataFrame with the reference's warning when nothing was extracted (:98-100).
    No SMILExtract binary is run: `opensmile_exe_path` is accepted and ignored, `config_file_path` (when it exists) supplies the
    frame / pre-emphasis / mel / MFCC / smoothing / delta settings, and the descriptors come from mshds_lld_extract -- the 720
    of the 911 Androids.conf columns built so far (components that are not built are named once when `verbose`)."""
    import pandas as pd

    params = {}
    if config_file_path and os.path.exists(config_file_path):
        params, missing = parse_smile_config(config_file_path)
        if verbose and missing:
            print(f"Note: components of '{os.path.basename(config_file_path)}' that this extractor does not build: {', '.join(missing)}")
    df = extract_lld_functionals(input_df, audio_file_column=audio_file_column, verbose=verbose, device=device, **params)
    ok = ~df.iloc[:, 1:].isna().all(axis=1)
    df = df[ok].reset_index(drop=True)
    if df.empty:
        print("Warning: No features were successfully extracted. The returned DataFrame is empty.")
        return pd.DataFrame()
    return df[[c for c in df.columns if c != "filename"] + ["filename"]]
