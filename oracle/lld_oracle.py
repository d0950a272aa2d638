"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the OpenSMILE low-level-descriptor slice (never imported by the product).

Follows the component chain of /root/reference/Androids.conf (cFramer :73-78, cVectorPreemphasis :80-83, cWindower :85-89,
cTransformFFT / cFFTmagphase :93-99, cMelspec :101-107, cMfcc :109-115, cEnergy :117-123, cMZcr :125-132) and the mean /
stddev functionals, with the definitions written out in include/mshds_b200.h.  PARITY UNPINNED: the SMILExtract 3.0.2 binary
the reference shells out to (src/opensmile_extractor.py:62-75) is not available offline and the repository holds no
OpenSMILE output; details such as the first pre-emphasised sample and the exact triangle evaluation are this restatement's
reading of the components.
"""
import numpy as np

DEFAULTS = dict(frame_size=0.025, frame_step=0.010, preemph=0.97, n_fft=0, n_mel=26, mel_lo=20.0, mel_hi=8000.0, n_mfcc=12,
                cep_lifter=22.0, smooth_win=3, delta_win=2)


def _mel(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def smooth_delta(rows: np.ndarray, smooth_win: int, delta_win: int) -> np.ndarray:
    """cContourSmoother (moving average, end frames repeated) then cDeltaRegression on the smoothed contours."""
    T = len(rows)
    if T == 0 or (smooth_win <= 1 and delta_win <= 0):
        return rows if delta_win <= 0 else np.zeros((0, 2 * rows.shape[1]))
    clip = lambda k: np.clip(k, 0, T - 1)
    t = np.arange(T)
    h = smooth_win // 2 if smooth_win > 1 else 0

    def sma(k):
        k = clip(k)
        acc = np.zeros((len(k), rows.shape[1]))
        for u in range(-h, h + 1):
            acc = acc + rows[clip(k + u)]
        return acc / (2 * h + 1)
    y = sma(t)
    if delta_win <= 0:
        return y
    acc = np.zeros_like(y)
    for i in range(1, delta_win + 1):
        acc = acc + i * (sma(t + i) - sma(t - i))
    d = acc / sum(2.0 * i * i for i in range(1, delta_win + 1))
    return np.concatenate([y, d], axis=1)


def frame_lld(x: np.ndarray, fs: float, **kw):
    """x float64 samples of ONE clip -> [n_frames, W] rows: (smoothed) mfcc 1..n, rms energy, zcr, then their deltas."""
    p = dict(DEFAULTS); p.update(kw)
    return smooth_delta(_raw_lld(x, fs, p), p["smooth_win"], p["delta_win"])


def _raw_lld(x: np.ndarray, fs: float, p: dict):
    nf = int(np.floor(p["frame_size"] * fs + 0.5)); ns = int(np.floor(p["frame_step"] * fs + 0.5))
    n_fft = p["n_fft"]
    if n_fft == 0:
        n_fft = 64
        while n_fft < nf:
            n_fft *= 2
    n_mel, n_mfcc, L, k = p["n_mel"], p["n_mfcc"], p["cep_lifter"], p["preemph"]
    nx = len(x)
    n_frames = (nx - nf) // ns + 1 if nx >= nf else 0
    D = n_mfcc + 2
    out = np.zeros((n_frames, D))
    if n_frames == 0:
        return out
    idx = np.arange(nf)[None, :] + ns * np.arange(n_frames)[:, None]
    fr = x[idx]                                                     # raw frames
    zcr = (fr[:, 1:] * fr[:, :-1] < 0).sum(axis=1) / (nf - 1) if nf > 1 else np.zeros(n_frames)
    pe = np.empty_like(fr)
    pe[:, 0] = fr[:, 0] * (1.0 - k)
    pe[:, 1:] = fr[:, 1:] - k * fr[:, :-1]
    w = 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(nf) / (nf - 1))
    xw = pe * w[None, :]
    energy = np.sqrt((xw * xw).sum(axis=1) / nf)
    mag = np.abs(np.fft.rfft(xw, n=n_fft, axis=1))                  # [n_frames, n_fft/2 + 1]
    hi = min(p["mel_hi"], 0.5 * fs)
    melbin = _mel(np.arange(n_fft // 2 + 1) * fs / n_fft)
    c = _mel(p["mel_lo"]) + (_mel(hi) - _mel(p["mel_lo"])) * np.arange(n_mel + 2) / (n_mel + 1)
    H = np.zeros((n_mel, n_fft // 2 + 1))
    for m in range(n_mel):
        up = (melbin - c[m]) / (c[m + 1] - c[m]); down = (c[m + 2] - melbin) / (c[m + 2] - c[m + 1])
        H[m] = np.maximum(0.0, np.minimum(up, down))
    E = mag @ H.T
    lm = np.log(np.maximum(E, 1e-10))
    i = np.arange(1, n_mfcc + 1)[:, None]; m = np.arange(n_mel)[None, :]
    dct = np.cos(np.pi * i * (m + 0.5) / n_mel)
    cc = np.sqrt(2.0 / n_mel) * (lm @ dct.T)
    if L > 0:
        cc = cc * (1.0 + 0.5 * L * np.sin(np.pi * np.arange(1, n_mfcc + 1) / L))[None, :]
    out[:, :n_mfcc] = cc
    out[:, n_mfcc] = energy
    out[:, n_mfcc + 1] = zcr
    return out


def extract(pcm: np.ndarray, offsets: np.ndarray, fs: float, **kw):
    """packed int16 batch -> (functionals [n, 2D], list of per-clip frame matrices)."""
    n = len(offsets) - 1
    rows = []
    D = ((kw.get("n_mfcc", DEFAULTS["n_mfcc"])) + 2) * (2 if kw.get("delta_win", DEFAULTS["delta_win"]) > 0 else 1)
    fun = np.full((n, 2 * D), np.nan)
    for i in range(n):
        x = pcm[offsets[i]:offsets[i + 1]].astype(np.float64) / 32768.0
        f = frame_lld(x, fs, **kw)
        rows.append(f)
        if len(f):
            fun[i, :D] = f.mean(axis=0)
            fun[i, D:] = f.std(axis=0)                              # population (ddof = 0)
    return fun, rows
