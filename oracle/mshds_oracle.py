"""oracle/mshds_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED.

ctypes front-end of the CPU oracle (oracle/*.c): a float64 restatement of the Praat 6.1.38 routines that
/root/reference/src/mshds_extractor.py reaches through praat-parselmouth 0.4.6 (un-vendored, not installable
here).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (robust_speech_analysis_framework_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmshds_oracle.so")

FEATURE_NAMES = [  # mshds_extractor.py:397-404
    'Speaking_Rate', 'Articulation_Rate', 'Phonation_Ratio', 'Pause_Rate', 'Mean_Pause_Duration',
    'mean_F0', 'stdev_F0_Semitone', 'mean_dB', 'range_ratio_dB', 'HNR_dB',
    'Spectral_Slope', 'Spectral_Tilt', 'Cepstral_Peak_Prominence',
    'mean_F1_Loc', 'std_F1_Loc', 'mean_B1_Loc', 'std_B1_Loc',
    'mean_F2_Loc', 'std_F2_Loc', 'mean_B2_Loc', 'std_B2_Loc',
    'Spectral_Gravity', 'Spectral_Std_Dev', 'Spectral_Skewness', 'Spectral_Kurtosis',
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force: bool = False) -> str:
    """Compile oracle/libmshds_oracle.so with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_interpolate_sinc.restype = C.c_double
        _lib.orc_improve_extremum.restype = C.c_double
        _lib.orc_bessel_i0.restype = C.c_double
        _lib.orc_quantile.restype = C.c_double
        _lib.orc_burg.restype = C.c_double
        _lib.orc_cpp.restype = C.c_double
        for name in ("orc_intensity", "orc_pitch", "orc_pulses", "orc_resample", "orc_formants", "orc_silences"):
            getattr(_lib, name).restype = C.c_long
    return _lib


OPTIONS = {   # SURVEY.md Appendix C switches (praat_core.h OrcOptions): name -> alternative values
    "silence_boundary": (1,), "cut_interval": (1,), "theil_tilt_complete": (1,), "theil_cpps_complete": (1,),
    "cpps_fit_range": (1,), "cpps_time_frames": (1,), "cpps_smooth_align": (1,), "vuv_overlap": (1, 2), "ltas_fill": (1,),
    "candidate_bound": (1,),
}


def set_option(name: str, value: int) -> None:
    """Flip one alternative reading of a Praat detail (oracle only; 0 = the default all goldens use)."""
    if not lib().orc_set_option(name.encode(), C.c_int(int(value))):
        raise KeyError(name)


def reset_options() -> None:
    lib().orc_set_option(b"reset", C.c_int(0))


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def pcm_to_float(pcm: np.ndarray) -> np.ndarray:
    return np.asarray(pcm, dtype=np.int16).astype(np.float64) / 32768.0


def extract(pcm: np.ndarray, offsets: np.ndarray, fs: float = 16000.0, nthreads: int = 1):
    """Full 25-column extraction of a packed int16 batch -> (features [n,25] f64, status [n] u32)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    n = len(offsets) - 1
    out = np.full((n, 25), np.nan)
    status = np.zeros(n, dtype=np.uint32)
    lib().orc_extract(pcm.ctypes.data_as(C.POINTER(C.c_int16)), offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                      C.c_int(n), C.c_double(fs), _p(out), status.ctypes.data_as(C.POINTER(C.c_uint32)),
                      C.c_int(nthreads))
    return out, status


def extract_f64(x: np.ndarray, fs: float = 16000.0):
    x = _f64(x)
    out = np.full(25, np.nan)
    st = C.c_uint32(0)
    lib().orc_extract_f64(_p(x), C.c_long(len(x)), C.c_double(fs), _p(out), C.byref(st))
    return out, st.value


def frame_grid(nx: int, fs: float, window: float, step: float):
    n = C.c_long(0)
    t1 = C.c_double(0)
    ok = lib().orc_frame_grid(C.c_long(nx), C.c_double(fs), C.c_double(window), C.c_double(step), C.byref(n), C.byref(t1))
    return (n.value, t1.value) if ok else None


def intensity(x, fs, min_pitch, dt, subtract_mean=True):
    x = _f64(x)
    cap = int(len(x) / fs / max(dt, 1e-4)) + 16
    out = np.zeros(cap)
    x1 = C.c_double(0)
    nf = lib().orc_intensity(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(min_pitch), C.c_double(dt),
                             C.c_int(int(subtract_mean)), _p(out), C.c_long(cap), C.byref(x1))
    if nf < 0:
        return None
    return out[:nf].copy(), x1.value


def intensity_stats(x, fs, min_pitch, dt):
    x = _f64(x)
    out = np.zeros(4)
    ok = lib().orc_intensity_stats(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(min_pitch), C.c_double(dt), _p(out))
    return out if ok else None


def pitch(x, fs, method=0, dt=0.005, floor=75.0, ppw=3.0, maxc=15, sil=0.03, vt=0.45, octc=0.01, jump=0.35, vuv=0.14,
          ceiling=600.0):
    """Selected path of Sound_to_Pitch_any (method 0 = AC Hanning, 2 = forward cross-correlation)."""
    x = _f64(x)
    step = dt if dt > 0 else ppw / floor / 4.0
    cap = int(len(x) / fs / step) + 16
    f = np.zeros(cap)
    s = np.zeros(cap)
    nc = np.zeros(cap, dtype=np.int32)
    x1 = C.c_double(0)
    dto = C.c_double(0)
    nf = lib().orc_pitch(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_int(method), C.c_double(dt), C.c_double(floor),
                         C.c_double(ppw), C.c_int(maxc), C.c_double(sil), C.c_double(vt), C.c_double(octc),
                         C.c_double(jump), C.c_double(vuv), C.c_double(ceiling), _p(f), _p(s),
                         nc.ctypes.data_as(_ip), C.c_long(cap), C.byref(x1), C.byref(dto))
    if nf < 0:
        return None
    return dict(freq=f[:nf].copy(), strength=s[:nf].copy(), ncand=nc[:nf].copy(), x1=x1.value, dt=dto.value)


def pulses(x, fs, method=0, dt=0.005, floor=75.0, ppw=3.0, vt=0.45, ceiling=600.0):
    x = _f64(x)
    cap = int(len(x) / fs * ceiling) + 64
    t = np.zeros(cap)
    n = lib().orc_pulses(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_int(method), C.c_double(dt), C.c_double(floor),
                         C.c_double(ppw), C.c_double(vt), C.c_double(ceiling), _p(t), C.c_long(cap))
    return None if n < 0 else t[:n].copy()


def pitch_values(x, fs=16000.0):
    x = _f64(x)
    fl = C.c_double(0)
    ce = C.c_double(0)
    fb = lib().orc_pitch_values(_p(x), C.c_long(len(x)), C.c_double(fs), C.byref(fl), C.byref(ce))
    return fl.value, ce.value, bool(fb)


def speechrate(x, fs=16000.0):
    x = _f64(x)
    out = np.full(5, np.nan)
    ok = lib().orc_speechrate(_p(x), C.c_long(len(x)), C.c_double(fs), _p(out))
    return out, bool(ok)


def hnr(x, fs, dt=0.005, floor=75.0, sil=0.1, ppw=4.5):
    x = _f64(x)
    m = C.c_double(0)
    ok = lib().orc_hnr(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(dt), C.c_double(floor), C.c_double(sil),
                       C.c_double(ppw), C.byref(m))
    return m.value if ok else None


def ltas(x, fs, floor, ceiling):
    x = _f64(x)
    bands = np.zeros(64)
    out2 = np.full(2, np.nan)
    ok = lib().orc_ltas(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(floor), C.c_double(ceiling), _p(bands), _p(out2))
    return (bands[:50].copy(), out2) if ok else None


def resample(x, fs, new_fs, precision):
    x = _f64(x)
    cap = int(len(x) * new_fs / fs) + 16
    out = np.zeros(cap)
    x1 = C.c_double(0)
    n = lib().orc_resample(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(new_fs), C.c_long(precision), _p(out),
                           C.c_long(cap), C.byref(x1))
    return None if n < 0 else (out[:n].copy(), x1.value)


def formants(x, fs):
    x = _f64(x)
    cap = int(len(x) / fs / 0.005) + 16
    f = np.zeros((cap, 5))
    bw = np.zeros((cap, 5))
    nf = np.zeros(cap, dtype=np.int32)
    x1 = C.c_double(0)
    n = lib().orc_formants(_p(x), C.c_long(len(x)), C.c_double(fs), _p(f), _p(bw), nf.ctypes.data_as(_ip), C.c_long(cap),
                           C.byref(x1))
    return None if n < 0 else dict(f=f[:n].copy(), bw=bw[:n].copy(), n=nf[:n].copy(), x1=x1.value)


def formant_stats(x, fs, floor, ceiling):
    x = _f64(x)
    out = np.full(8, np.nan)
    lib().orc_formant_stats(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(floor), C.c_double(ceiling), _p(out))
    return out


def cpp(x, fs, floor, ceiling):
    x = _f64(x)
    return lib().orc_cpp(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(floor), C.c_double(ceiling))


def cpps_segment(x, fs):
    x = _f64(x)
    v = C.c_double(0)
    ok = lib().orc_cpps_segment(_p(x), C.c_long(len(x)), C.c_double(fs), C.byref(v))
    return v.value if ok else None


def moments(x, fs, floor, ceiling):
    x = _f64(x)
    out = np.full(4, np.nan)
    ok = lib().orc_moments(_p(x), C.c_long(len(x)), C.c_double(fs), C.c_double(floor), C.c_double(ceiling), _p(out))
    return out if ok else None


# ---- numerics -------------------------------------------------------------------------------------------------

def interpolate_sinc(y, x, depth):
    y = _f64(y)
    return lib().orc_interpolate_sinc(_p(y), C.c_long(len(y)), C.c_double(x), C.c_long(depth))


def improve_extremum(y, ixmid, interpolation, is_maximum=True):
    y = _f64(y)
    xr = C.c_double(0)
    v = lib().orc_improve_extremum(_p(y), C.c_long(len(y)), C.c_long(ixmid), C.c_int(interpolation),
                                   C.c_int(int(is_maximum)), C.byref(xr))
    return v, xr.value


def bessel_i0(x):
    return lib().orc_bessel_i0(C.c_double(x))


def quantile(sorted_values, q):
    a = _f64(sorted_values)
    return lib().orc_quantile(_p(a), C.c_long(len(a)), C.c_double(q))


def theil(x, y, complete=False):
    x = _f64(x)
    y = _f64(y)
    m = C.c_double(0)
    b = C.c_double(0)
    lib().orc_theil(_p(x), _p(y), C.c_long(len(x)), C.c_int(int(complete)), C.byref(m), C.byref(b))
    return m.value, b.value


def burg(x, order):
    x = _f64(x)
    a = np.zeros(order)
    xms = lib().orc_burg(_p(x), C.c_long(len(x)), C.c_int(order), _p(a))
    return a, xms


def roots(coeffs_ascending):
    c = _f64(coeffs_ascending)
    n = len(c) - 1
    re = np.zeros(n)
    im = np.zeros(n)
    k = lib().orc_roots(_p(c), C.c_int(n), _p(re), _p(im))
    return re[:k] + 1j * im[:k]


def fft(re, im, sign=-1):
    re = _f64(re).copy()
    im = _f64(im).copy()
    lib().orc_fft(_p(re), _p(im), C.c_long(len(re)), C.c_int(sign))
    return re + 1j * im


def silences(contour, dx, x1, xmin, xmax, thr, min_sil, min_snd):
    c = _f64(contour)
    cap = len(c) + 2
    b = np.zeros((cap, 2))
    s = np.zeros(cap, dtype=np.int32)
    n = lib().orc_silences(_p(c), C.c_long(len(c)), C.c_double(dx), C.c_double(x1), C.c_double(xmin), C.c_double(xmax),
                           C.c_double(thr), C.c_double(min_sil), C.c_double(min_snd), _p(b), s.ctypes.data_as(_ip),
                           C.c_long(cap))
    return b[:n].copy(), s[:n].astype(bool)
