"""Host-side mirror of the reference's MSHDS extractor interface.

Reference: /root/reference/src/mshds_extractor.py
  * extract_mshds_features(input_df, audio_file_column='filepath', verbose=True) -> DataFrame   (:379-459)
    one row per input row, in input order, columns ['filename'] + the 25 names of :397-404, float64 values;
    it never raises: a helper failure gives NaN for that helper's columns (:124,161,182,204,224,250,300,337,375) and a
    file that cannot be loaded gives a NaN row with the filename kept and, iff verbose, the printed message of :452-453.

Here the per-file Praat calls are replaced by ONE batched call into libmshds_b200.so (include/mshds_b200.h) per group
of recordings; WAV decoding and mono mix-down stay on the host (the reference does them inside parselmouth.Sound, :415-417).
There is no CPU fallback: without the CUDA library / a CUDA device the function raises at the first call.
"""
from __future__ import annotations

import os
import wave

import numpy as np

from . import _lib

FEATURE_NAMES = list(_lib.FEATURE_NAMES)
TARGET_RATE = 16000


class AudioLoadError(Exception):
    pass


def read_wav_mono_int16(path: str):
    """Decode a PCM WAV into (int16 mono samples, sample_rate).

    16-bit files are passed through bit-exactly (what parselmouth.Sound(path) holds is sample/32768, :415); multi-channel
    audio is averaged like Sound.convert_to_mono (:416-417).  8/24/32-bit PCM is re-quantised to int16.
    """
    try:
        with wave.open(path, "rb") as w:
            nch, sw, fs, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
    except Exception as e:  # includes FileNotFoundError, wave.Error, EOFError
        raise AudioLoadError(str(e)) from e
    if sw == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.int32)
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.int32) - 128) << 8
    elif sw == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        x = ((b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)) << 8 >> 8) >> 8
    elif sw == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.int64) >> 16
    else:
        raise AudioLoadError(f"unsupported sample width {sw}")
    if nch > 1:
        x = x.reshape(-1, nch)
        x = np.floor(x.mean(axis=1) + 0.5)
    return np.clip(x, -32768, 32767).astype(np.int16), int(fs)


def read_wav_mono(path: str):
    """Decode a PCM WAV into (mono samples, sample_rate) the way parselmouth.Sound(path) + convert_to_mono hold it (:415-417).

    Mono 16-bit audio stays int16 (value / 32768 is formed on the device, exactly); everything else -- 8/24/32-bit samples,
    several channels (Praat averages the channels in floating point, which int16 cannot hold) -- is returned as float64 in
    [-1, 1) and goes through the library's float64 entry (MSHDS_PCM_FLOAT64).
    """
    try:
        with wave.open(path, "rb") as w:
            nch, sw, fs, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
    except Exception as e:
        raise AudioLoadError(str(e)) from e
    if sw == 2:
        x = np.frombuffer(raw, dtype="<i2")
        if nch == 1:
            return x.copy(), int(fs)
        x = x.astype(np.float64) / 32768.0
    elif sw == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float64) - 128.0) / 128.0
    elif sw == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        x = (((b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)) << 8) >> 8).astype(np.float64) / 8388608.0
    elif sw == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0
    else:
        raise AudioLoadError(f"unsupported sample width {sw}")
    if nch > 1:
        x = x.reshape(-1, nch).mean(axis=1)
    return np.ascontiguousarray(x, dtype=np.float64), int(fs)


_EXTRACTORS = {}


def get_extractor(device: int = 0) -> "_lib.Extractor":
    ex = _EXTRACTORS.get(device)
    if ex is None:
        ex = _lib.Extractor(device)
        _EXTRACTORS[device] = ex
    return ex


def extract_mshds_from_pcm(pcm: np.ndarray, offsets: np.ndarray, sample_rate: int = TARGET_RATE, device: int = 0):
    """Tensor-level entry: packed int16 batch + offsets -> (features [n,25] float64, status [n] uint32)."""
    return get_extractor(device).extract_host(pcm, offsets, sample_rate)


def extract_mshds_features(input_df, audio_file_column='filepath', verbose=True, device: int = 0, max_batch_seconds: float = 7200.0):
    """Drop-in for /root/reference/src/mshds_extractor.py:379 (same name, arguments, columns, NaN conventions).

    `device` and `max_batch_seconds` are additions with defaults; every recording is still processed independently.
    """
    import pandas as pd

    ex = get_extractor(device)
    paths = [row[audio_file_column] for _, row in input_df.iterrows()]
    filenames = [os.path.basename(p) for p in paths]
    n = len(paths)
    feats = np.full((n, len(FEATURE_NAMES)), np.nan, dtype=np.float64)

    # one open batch per sampling frequency: a device call takes clips of one rate (the library resamples to 16 kHz itself,
    # the reference's snd.resample(16000, 50) of :418-419)
    batches = {}

    def flush(fs):
        batch_idx, batch_pcm, _ = batches.pop(fs, ([], [], 0))
        if not batch_idx:
            return
        offs = np.cumsum([0] + [len(p) for p in batch_pcm]).astype(np.int64)
        pcm = np.concatenate(batch_pcm)
        try:
            if pcm.dtype == np.int16:
                out, _status = ex.extract_host(pcm, offs, fs[0])
            else:
                out, _status = ex.extract_host_f64(pcm, offs, fs[0])
            feats[np.asarray(batch_idx)] = out
        except Exception as e:  # mirrors the whole-file handler at :450-457
            if verbose:
                for i in batch_idx:
                    print(f"ERROR processing file '{filenames[i]}': {e}. Appending NaNs.")

    iterator = range(n)
    if verbose:
        try:
            from tqdm.auto import tqdm
            iterator = tqdm(iterator, total=n, desc="Extracting MSHDS Features")
        except Exception:
            pass
    for i in iterator:
        try:
            pcm, rate = read_wav_mono(paths[i])
            fs = (rate, pcm.dtype == np.int16)            # one open batch per (rate, sample type)
            if len(pcm) == 0:
                raise AudioLoadError("empty sound")
        except Exception as e:
            if verbose:
                print(f"ERROR processing file '{filenames[i]}': {e}. Appending NaNs.")
            continue
        b = batches.setdefault(fs, ([], [], 0))
        b[0].append(i)
        b[1].append(pcm)
        batches[fs] = (b[0], b[1], b[2] + len(pcm))
        if batches[fs][2] >= max_batch_seconds * rate:
            flush(fs)
    for fs in list(batches):
        flush(fs)

    rows = []
    for i in range(n):
        d = {'filename': filenames[i]}
        for k, name in enumerate(FEATURE_NAMES):
            d[name] = float(feats[i, k])
        rows.append(d)
    return pd.DataFrame(rows)
