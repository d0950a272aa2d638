"""ctypes binding of libmshds_b200.so (C ABI: include/mshds_b200.h).

There is deliberately NO fallback: if the CUDA library cannot be built/loaded or no CUDA device is present, every
compute entry point raises.  The CPU oracle under oracle/ is test infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

FEATURE_NAMES = [  # /root/reference/src/mshds_extractor.py:397-404
    'Speaking_Rate', 'Articulation_Rate', 'Phonation_Ratio', 'Pause_Rate', 'Mean_Pause_Duration',
    'mean_F0', 'stdev_F0_Semitone', 'mean_dB', 'range_ratio_dB', 'HNR_dB',
    'Spectral_Slope', 'Spectral_Tilt', 'Cepstral_Peak_Prominence',
    'mean_F1_Loc', 'std_F1_Loc', 'mean_B1_Loc', 'std_B1_Loc',
    'mean_F2_Loc', 'std_F2_Loc', 'mean_B2_Loc', 'std_B2_Loc',
    'Spectral_Gravity', 'Spectral_Std_Dev', 'Spectral_Skewness', 'Spectral_Kurtosis',
]
N_FEATURES = 25
PCM_ON_DEVICE = 1
OUT_ON_DEVICE = 2
PCM_FLOAT64 = 4
AGG_ON_DEVICE = 1
CONTOURS = {"f0": 0, "intensity": 1, "hnr": 2, "formants": 3, "moments": 4}

EXPORTED_SYMBOLS = [
    "mshds_create", "mshds_destroy", "mshds_set_stream", "mshds_set_chunk_samples", "mshds_last_error", "mshds_extract",
    "mshds_launch_count", "mshds_debug_fetch", "mshds_profile_enable", "mshds_profile_report", "mshds_aggregate_sessions",
    "mshds_lld_default_params", "mshds_lld_extract", "mshds_extract_contours", "mshds_reset_stream", "mshds_set_option", "mshds_fp64_peak",
]

_lib = None


class MshdsError(RuntimeError):
    pass


class LldParams(C.Structure):
    """struct mshds_lld_params (include/mshds_b200.h)."""
    _fields_ = [("frame_size", C.c_double), ("frame_step", C.c_double), ("preemph", C.c_double), ("n_fft", C.c_int),
                ("n_mel", C.c_int), ("mel_lo", C.c_double), ("mel_hi", C.c_double), ("n_mfcc", C.c_int),
                ("cep_lifter", C.c_double), ("smooth_win", C.c_int), ("delta_win", C.c_int),
                ("descriptor_set", C.c_int), ("functional_set", C.c_int)]


def load(build_if_needed: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_needed and _build.needs_build():
        _build.build()
    if not os.path.exists(path):
        raise MshdsError(f"{path} is missing: build it with `python -m robust_speech_analysis_framework_b200.build`")
    lib = C.CDLL(path)
    lib.mshds_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.mshds_destroy.argtypes = [C.c_void_p]
    lib.mshds_destroy.restype = None
    lib.mshds_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.mshds_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.mshds_reset_stream.argtypes = [C.c_void_p]
    lib.mshds_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
    lib.mshds_set_chunk_samples.argtypes = [C.c_void_p, C.c_longlong]
    lib.mshds_last_error.argtypes = [C.c_void_p]
    lib.mshds_last_error.restype = C.c_char_p
    lib.mshds_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint]
    lib.mshds_launch_count.argtypes = [C.c_void_p]
    lib.mshds_launch_count.restype = C.c_longlong
    lib.mshds_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.mshds_profile_report.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.mshds_aggregate_sessions.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_uint]
    lib.mshds_lld_default_params.argtypes = [C.POINTER(LldParams)]
    lib.mshds_lld_default_params.restype = None
    lib.mshds_lld_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(LldParams), C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_uint]
    lib.mshds_extract_contours.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                           C.c_void_p, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_uint]
    lib.mshds_debug_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    _lib = lib
    return lib


class Extractor:
    """One handle = one CUDA device + stream + scratch (mshds_create / mshds_destroy)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.mshds_create(int(device), C.byref(h))
        if rc != 0 or not h:
            raise MshdsError(f"mshds_create(device={device}) failed with code {rc}: no usable CUDA device "
                             "(this package has no CPU fallback)")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mshds_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise MshdsError(f"libmshds_b200 error {rc}: {self._lib.mshds_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream_ptr: int | None):
        """Run on the caller's stream; 0 / None is the legacy default stream (torch's default), not the private one."""
        self._check(self._lib.mshds_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    def reset_stream(self):
        self._check(self._lib.mshds_reset_stream(self._h))

    def set_option(self, name: str, value: int):
        self._check(self._lib.mshds_set_option(self._h, name.encode(), int(value)))

    def set_chunk_samples(self, n: int):
        self._check(self._lib.mshds_set_chunk_samples(self._h, int(n)))

    @property
    def launch_count(self) -> int:
        return int(self._lib.mshds_launch_count(self._h))

    def extract_host(self, pcm: np.ndarray, offsets: np.ndarray, sample_rate: int = 16000):
        """Host int16 buffer in, host float64 [n,25] + uint32 [n] out (H2D and D2H inside the call)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        out = np.full((n, N_FEATURES), np.nan, dtype=np.float64)
        status = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.mshds_extract(self._h, pcm.ctypes.data, offsets.ctypes.data, n, int(sample_rate),
                                            out.ctypes.data, status.ctypes.data, 0))
        return out, status

    def extract_host_f64(self, samples: np.ndarray, offsets: np.ndarray, sample_rate: int = 16000):
        """Same as extract_host for float64 samples in [-1, 1) (MSHDS_PCM_FLOAT64): 24/32-bit, float and multi-channel files."""
        samples = np.ascontiguousarray(samples, dtype=np.float64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        out = np.full((max(n, 0), N_FEATURES), np.nan, dtype=np.float64)
        status = np.zeros(max(n, 0), dtype=np.uint32)
        self._check(self._lib.mshds_extract(self._h, samples.ctypes.data, offsets.ctypes.data, n, int(sample_rate),
                                            out.ctypes.data, status.ctypes.data, PCM_FLOAT64))
        return out, status

    def extract_device(self, pcm_ptr: int, offsets: np.ndarray, out_ptr: int, status_ptr: int, sample_rate: int = 16000):
        """Device-resident int16 batch in, device float64 [n,25] / uint32 [n] out (raw device pointers)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        self._check(self._lib.mshds_extract(self._h, C.c_void_p(pcm_ptr), offsets.ctypes.data, n, int(sample_rate),
                                            C.c_void_p(out_ptr), C.c_void_p(status_ptr), PCM_ON_DEVICE | OUT_ON_DEVICE))

    def fp64_peak_tflops(self) -> float:
        """Measured DFMA issue peak of this device (mshds_fp64_peak)."""
        v = C.c_double(0.0)
        self._check(self._lib.mshds_fp64_peak(self._h, C.byref(v)))
        return float(v.value)

    def profile(self, on: bool):
        self._check(self._lib.mshds_profile_enable(self._h, int(on)))

    def extract_contours(self, pcm: np.ndarray, offsets: np.ndarray, which: str, sample_rate: int = 16000):
        """mshds_extract_contours: per-frame contour `which` in CONTOURS for every clip.
        Returns dict(values [rows, width], frame_offsets [n + 1], t1 [n], dt, features [n, 25])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        lens = np.diff(offsets).astype(np.float64)
        cap = int(np.sum(np.floor(lens / sample_rate / 0.005) + 2)) + 1
        width = C.c_int(0)
        dt = C.c_double(0.0)
        values = np.full((cap, 4), np.nan)
        fo = np.zeros(n + 1, dtype=np.int64)
        t1 = np.full(max(n, 1), np.nan)
        feats = np.full((max(n, 0), N_FEATURES), np.nan)
        # the library writes rows of `width` doubles; width <= 4, so a [cap, 4] buffer always suffices
        self._check(self._lib.mshds_extract_contours(self._h, pcm.ctypes.data, offsets.ctypes.data, n, int(sample_rate), CONTOURS[which],
                                                     values.ctypes.data, cap, fo.ctypes.data, t1.ctypes.data, C.byref(dt), C.byref(width),
                                                     feats.ctypes.data if n > 0 else None, 0))
        rows, w = int(fo[-1]), width.value
        vals = values.ravel()[: rows * w].reshape(rows, w).copy()
        return dict(values=vals, frame_offsets=fo, t1=t1[:n], dt=dt.value, features=feats)

    def lld_params(self, **kw) -> "LldParams":
        p = LldParams()
        self._lib.mshds_lld_default_params(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise MshdsError(f"unknown LLD parameter {k!r}")
            setattr(p, k, v)
        return p

    def lld_extract(self, pcm: np.ndarray, offsets: np.ndarray, sample_rate: int = 16000, want_frames: bool = False, **params):
        """mshds_lld_extract on host arrays -> (functionals [n, NF * D], frames [total, D] or None, frame_offsets [n + 1]);
        D = width of a frame row, NF = 2 (amean, stddev) or 12 (functional_set = 1), functional-major."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        p = self.lld_params(**params)
        D = (p.n_mfcc + 2 + (16 if p.descriptor_set else 0)) * (2 if p.delta_win > 0 else 1)           # width of a frame row
        fun = np.full((max(n, 0), (12 if p.functional_set else 2) * D), np.nan)
        fo = np.zeros(n + 1, dtype=np.int64)
        frames = None
        if want_frames:
            nf, ns = int(np.floor(p.frame_size * sample_rate + 0.5)), int(np.floor(p.frame_step * sample_rate + 0.5))
            lens = np.diff(offsets)
            total = int(np.where(lens >= nf, (lens - nf) // max(ns, 1) + 1, 0).sum()) if nf >= 2 and ns >= 1 else 0
            frames = np.zeros((total, D))
        self._check(self._lib.mshds_lld_extract(self._h, pcm.ctypes.data, offsets.ctypes.data, n, int(sample_rate), C.byref(p),
                                                fun.ctypes.data, frames.ctypes.data if frames is not None and frames.size else None,
                                                fo.ctypes.data, 0))
        return fun, frames, fo

    def lld_extract_device(self, pcm_ptr: int, offsets: np.ndarray, out_ptr: int, sample_rate: int = 16000, **params):
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        p = self.lld_params(**params)
        self._check(self._lib.mshds_lld_extract(self._h, C.c_void_p(pcm_ptr), offsets.ctypes.data, len(offsets) - 1, int(sample_rate),
                                                C.byref(p), C.c_void_p(out_ptr), None, None, PCM_ON_DEVICE | OUT_ON_DEVICE))

    def aggregate_sessions(self, features: np.ndarray, row_group: np.ndarray, n_groups: int):
        """mshds_aggregate_sessions on host arrays: (mean [g, d], std [g, d]) of the rows of every session."""
        features = np.ascontiguousarray(features, dtype=np.float64)
        if features.ndim != 2:
            raise MshdsError("features must be a 2-D array")
        row_group = np.ascontiguousarray(row_group, dtype=np.int32)
        n, d = features.shape
        if len(row_group) != n:
            raise MshdsError("row_group must have one entry per row")
        mean = np.full((n_groups, d), np.nan)
        std = np.full((n_groups, d), np.nan)
        self._check(self._lib.mshds_aggregate_sessions(self._h, features.ctypes.data, n, d, row_group.ctypes.data, int(n_groups),
                                                       mean.ctypes.data, std.ctypes.data, 0))
        return mean, std

    def aggregate_sessions_device(self, feat_ptr: int, n_rows: int, n_cols: int, row_group: np.ndarray, n_groups: int,
                                  mean_ptr: int, std_ptr: int):
        """Same on device buffers (e.g. straight from extract_device): nothing but the group list crosses PCIe."""
        row_group = np.ascontiguousarray(row_group, dtype=np.int32)
        self._check(self._lib.mshds_aggregate_sessions(self._h, C.c_void_p(feat_ptr), int(n_rows), int(n_cols), row_group.ctypes.data,
                                                       int(n_groups), C.c_void_p(mean_ptr), C.c_void_p(std_ptr), AGG_ON_DEVICE))

    def profile_report(self) -> dict:
        """{stage: (milliseconds, spans)} accumulated since profile(True)."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self._lib.mshds_profile_report(self._h, buf, len(buf)))
        rep = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.split("\t")
            rep[name] = (float(ms), int(cnt))
        return rep

    def debug_fetch(self, name: str, clip: int, dtype=np.float64, cap: int = 1 << 22) -> np.ndarray:
        buf = np.zeros(cap, dtype=dtype)
        n = C.c_size_t(0)
        self._check(self._lib.mshds_debug_fetch(self._h, name.encode(), int(clip), buf.ctypes.data, cap, C.byref(n)))
        if n.value > cap:
            return self.debug_fetch(name, clip, dtype, n.value)
        return buf[: n.value].copy()
