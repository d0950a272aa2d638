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
import threading
import warnings
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
    except wave.Error as e:
        # not integer-PCM WAV: IEEE-float / WAVE_FORMAT_EXTENSIBLE WAV and AIFF are decoded below; parselmouth.Sound(path) at
        # :415 also opens FLAC, MP3, NIST ... for which this image has no decoder -> a distinct, explicit error
        return _read_other_container(path, str(e))
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


def _read_other_container(path: str, wave_error: str):
    """IEEE-float / extensible WAV (scipy.io.wavfile) and AIFF (stdlib aifc) -> (float64 mono in [-1, 1), rate)."""
    with open(path, "rb") as f:
        magic = f.read(12)
    if magic[:4] == b"RIFF" and magic[8:12] == b"WAVE":
        try:
            from scipy.io import wavfile
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                fs, x = wavfile.read(path)
        except Exception as e:
            raise AudioLoadError(f"WAV container not readable ({wave_error}; scipy: {e})") from e
        if x.dtype.kind == "f":
            y = x.astype(np.float64)
        elif x.dtype == np.uint8:
            y = (x.astype(np.float64) - 128.0) / 128.0
        else:
            y = x.astype(np.float64) / float(1 << (8 * x.dtype.itemsize - 1))
        if y.ndim > 1:
            y = y.mean(axis=1)
        return np.ascontiguousarray(y), int(fs)
    if magic[:4] == b"FORM" and magic[8:12] in (b"AIFF", b"AIFC"):
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                import aifc
                with aifc.open(path, "rb") as a:
                    nch, sw, fs, n = a.getnchannels(), a.getsampwidth(), a.getframerate(), a.getnframes()
                    raw = a.readframes(n)
        except Exception as e:
            raise AudioLoadError(f"AIFF container not readable: {e}") from e
        if sw == 1:
            y = np.frombuffer(raw, dtype=np.int8).astype(np.float64) / 128.0
        elif sw == 2:
            y = np.frombuffer(raw, dtype=">i2").astype(np.float64) / 32768.0
        elif sw == 3:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            y = (((b[:, 2] | (b[:, 1] << 8) | (b[:, 0] << 16)) << 8) >> 8).astype(np.float64) / 8388608.0
        else:
            y = np.frombuffer(raw, dtype=">i4").astype(np.float64) / 2147483648.0
        if nch > 1:
            y = y.reshape(-1, nch).mean(axis=1)
        return np.ascontiguousarray(y), int(fs)
    kind = {b"fLaC": "FLAC", b"OggS": "Ogg", b"ID3": "MP3"}.get(magic[:4], {b"ID3": "MP3"}.get(magic[:3], "unknown"))
    # FLAC / Ogg / MP3 / NIST ...: decoded through the optional `soundfile` package (libsndfile) when the installation has it --
    # float64 frames, channels averaged in floating point like Praat's convert_to_mono (:416-417)
    try:
        import soundfile
    except ImportError:
        soundfile = None
    if soundfile is not None:
        try:
            y, fs = soundfile.read(path, dtype="float64", always_2d=True)
        except Exception as e:
            raise AudioLoadError(f"audio container ({kind}) not readable by soundfile: {e}") from e
        return np.ascontiguousarray(y.mean(axis=1)), int(fs)
    raise AudioLoadError(f"unsupported audio container ({kind}): parselmouth.Sound reads it, this drop-in decodes WAV (PCM / float) "
                         f"and AIFF only (plus whatever the optional `soundfile` package reads, which is not installed) -- convert the file to WAV")


_TLS = threading.local()
_LOCK = threading.Lock()


def get_extractor(device: int = 0) -> "_lib.Extractor":
    """One handle per (host thread, device): the C ABI is one-handle-per-thread (include/mshds_b200.h)."""
    cache = getattr(_TLS, "extractors", None)
    if cache is None:
        cache = _TLS.extractors = {}
    ex = cache.get(device)
    if ex is None:
        with _LOCK:                       # library load / first CUDA context creation, once at a time
            ex = _lib.Extractor(device)
        cache[device] = ex
    return ex


def extract_mshds_from_pcm(pcm: np.ndarray, offsets: np.ndarray, sample_rate: int = TARGET_RATE, device: int = 0):
    """Tensor-level entry: packed int16 batch + offsets -> (features [n,25] float64, status [n] uint32)."""
    return get_extractor(device).extract_host(pcm, offsets, sample_rate)


def _run_batch(device: int, pcm_list, rate: int, is_int16: bool):
    """One device call for a list of clips -> features [m, 25]; raises _lib.MshdsError on a device-side failure."""
    ex = get_extractor(device)
    offs = np.cumsum([0] + [len(p) for p in pcm_list]).astype(np.int64)
    pcm = np.concatenate(pcm_list)
    out, _status = ex.extract_host(pcm, offs, rate) if is_int16 else ex.extract_host_f64(pcm, offs, rate)
    return out


def _run_isolated(device: int, pcm_list, rate: int, is_int16: bool, names, verbose: bool):
    """A batch whose device call failed is retried one recording at a time, so that -- like the per-file try/except of the
    reference (:450-457) -- only the recording that cannot be processed becomes a NaN row.  Device errors are reported with
    warnings.warn even when verbose is off: they are environment failures (out of memory, lost context), not data."""
    try:
        return _run_batch(device, pcm_list, rate, is_int16)
    except Exception as e:
        warnings.warn(f"MSHDS device call failed for a batch of {len(pcm_list)} recordings on cuda:{device} ({e}); "
                      f"retrying one recording at a time", RuntimeWarning)
    out = np.full((len(pcm_list), len(FEATURE_NAMES)), np.nan)
    for k, clip in enumerate(pcm_list):
        try:
            out[k] = _run_batch(device, [clip], rate, is_int16)[0]
        except Exception as e:
            warnings.warn(f"MSHDS extraction failed for '{names[k]}' on cuda:{device}: {e}", RuntimeWarning)
            if verbose:
                print(f"ERROR processing file '{names[k]}': {e}. Appending NaNs.")
    return out


def extract_mshds_features(input_df, audio_file_column='filepath', verbose=True, device: int = 0, max_batch_seconds: float = 7200.0,
                           devices=None):
    """Drop-in for /root/reference/src/mshds_extractor.py:379 (same name, arguments, columns, NaN conventions).

    `device`, `devices` and `max_batch_seconds` are additions with defaults; every recording is still processed independently.
    `devices=[0, 1, ...]` spreads each batch over several GPUs of the box: clips are assigned by greedy longest-processing-time
    (sharding.lpt_assign), one host thread and one library handle per GPU, rows come back in input order.
    """
    import pandas as pd
    from .sharding import lpt_assign

    devs = [int(d) for d in devices] if devices else [int(device)]
    for d in devs:
        get_extractor(d)                 # fail loudly (no CPU fallback) before any file is read
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
        rate, is_int16 = fs
        if len(devs) == 1:
            feats[np.asarray(batch_idx)] = _run_isolated(devs[0], batch_pcm, rate, is_int16, [filenames[i] for i in batch_idx], verbose)
            return
        parts = lpt_assign([len(p) for p in batch_pcm], len(devs))
        results = [None] * len(devs)

        def work(r):
            if parts[r]:
                results[r] = _run_isolated(devs[r], [batch_pcm[k] for k in parts[r]], rate, is_int16,
                                           [filenames[batch_idx[k]] for k in parts[r]], verbose)

        threads = [threading.Thread(target=work, args=(r,)) for r in range(len(devs))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for r, part in enumerate(parts):
            if part:
                feats[np.asarray([batch_idx[k] for k in part])] = results[r]

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
        if batches[fs][2] >= max_batch_seconds * rate * len(devs):
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
