"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the OpenSMILE low-level-descriptor slice (never imported by the product).

Follows the component chain of /root/reference/Androids.conf (cFramer :73-78, cVectorPreemphasis :80-83, cWindower :85-89,
cTransformFFT / cFFTmagphase :93-99, cMelspec :101-107, cMfcc :109-115, cEnergy :117-123, cMZcr :125-132; with
descriptor_set = 1 also cIntensity :134-140 and 14 of the 16 cSpectral descriptors :257-282) and the functionals (mean / stddev,
or with functional_set = 1 the twelve of functL1 :349-366), with the definitions written out in include/mshds_b200.h.  PARITY UNPINNED: the SMILExtract 3.0.2 binary
the reference shells out to (src/opensmile_extractor.py:62-75) is not available offline and the repository holds no
OpenSMILE output; details such as the first pre-emphasised sample and the exact triangle evaluation are this restatement's
reading of the components.
"""
import numpy as np

DEFAULTS = dict(frame_size=0.025, frame_step=0.010, preemph=0.97, n_fft=0, n_mel=26, mel_lo=20.0, mel_hi=8000.0, n_mfcc=12,
                cep_lifter=22.0, smooth_win=3, delta_win=2, descriptor_set=0, functional_set=0)
BANDS = ((250.0, 650.0), (1000.0, 4000.0))            # Androids.conf:261-262
ROLLOFF = (0.25, 0.50, 0.75, 0.90)                    # Androids.conf:263-266


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
    D = n_mfcc + 2 + (16 if p["descriptor_set"] else 0)
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
    if p["descriptor_set"]:
        q = out[:, n_mfcc + 2:]
        Im = ((xw * xw) * w[None, :]).sum(axis=1) / w.sum() / 1e-6           # cIntensity: Hamming-weighted mean square / I0
        q[:, 0] = Im
        q[:, 1] = Im ** 0.3
        S = mag * mag
        N = n_fft // 2 + 1
        f = np.arange(N) * fs / n_fft
        for b, (lo, hi_) in enumerate(BANDS):
            q[:, 2 + b] = S[:, (f >= lo) & (f <= hi_)].sum(axis=1)
        tot = S.sum(axis=1)
        cum = np.cumsum(S, axis=1)
        for r, frac in enumerate(ROLLOFF):
            idx = np.array([np.argmax(cum[t] >= frac * cum[t, -1]) if (cum[t] >= frac * cum[t, -1]).any() else N - 1
                            for t in range(n_frames)])
            q[:, 4 + r] = f[idx]
        dm = np.diff(mag, axis=0)
        q[1:, 8] = np.sqrt((dm * dm).sum(axis=1) / N)
        ok = tot > 0
        st = np.where(ok, tot, 1.0)
        cen = np.where(ok, (S * f[None, :]).sum(axis=1) / st, 0.0)
        d = f[None, :] - cen[:, None]
        var = np.where(ok, (d ** 2 * S).sum(axis=1) / st, 0.0)
        sd = np.sqrt(var)
        vv = np.where(var > 0, var, 1.0)
        q[:, 9] = cen
        pk = S / st[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            ent = -np.where((pk > 0) & ok[:, None], pk * np.log2(pk), 0.0).sum(axis=1)
        q[:, 10] = ent
        q[:, 11] = var
        q[:, 12] = np.where(var > 0, (d ** 3 * S).sum(axis=1) / st / (vv * np.where(var > 0, sd, 1.0)), 0.0)
        q[:, 13] = np.where(var > 0, (d ** 4 * S).sum(axis=1) / st / (vv * vv), 0.0)
        sf, sff = f.sum(), (f * f).sum()
        den = N * sff - sf * sf
        q[:, 14] = (N * (S * f[None, :]).sum(axis=1) - sf * tot) / den if den != 0 else 0.0
        q[:, 15] = np.where(ok, np.exp(np.log(np.maximum(S, 1e-100)).mean(axis=1)) / (st / N), 0.0)
    return out


def functionals12(y: np.ndarray) -> np.ndarray:
    """[T, W] contours -> [12, W]: max, min, range, maxPos, minPos, amean, linregc1, linregc2, linregerrQ, stddev, skewness,
    kurtosis (Androids.conf functL1 :349-366; positions and regression abscissa in frames)."""
    T, W = y.shape
    t = np.arange(T, dtype=np.float64)
    mean = y.mean(axis=0)
    tbar = 0.5 * (T - 1.0)
    stt = T * (T * T - 1.0) / 12.0
    m = ((t[:, None] * y).sum(axis=0) - tbar * y.sum(axis=0)) / stt if T > 1 else np.zeros(W)
    b = mean - m * tbar
    e = y - mean[None, :]
    m2 = (e ** 2).mean(axis=0)
    sd = np.sqrt(m2)
    amax = np.maximum(np.abs(y.max(axis=0)), np.abs(y.min(axis=0)))
    shaped = m2 > 1e-24 * amax * amax                    # constant up to rounding: skewness = kurtosis = 0
    mm = np.where(shaped, m2, 1.0)
    res = y - (m[None, :] * t[:, None] + b[None, :])
    return np.stack([y.max(axis=0), y.min(axis=0), y.max(axis=0) - y.min(axis=0), y.argmax(axis=0).astype(np.float64),
                     y.argmin(axis=0).astype(np.float64), mean, m, b, (res ** 2).mean(axis=0), sd,
                     np.where(shaped, (e ** 3).mean(axis=0) / (mm * np.where(shaped, sd, 1.0)), 0.0),
                     np.where(shaped, (e ** 4).mean(axis=0) / (mm * mm), 0.0)])


def extract(pcm: np.ndarray, offsets: np.ndarray, fs: float, **kw):
    """packed int16 batch -> (functionals [n, NF * D], list of per-clip frame matrices); NF = 2 or 12, functional-major."""
    n = len(offsets) - 1
    rows = []
    D = ((kw.get("n_mfcc", DEFAULTS["n_mfcc"])) + 2 + (16 if kw.get("descriptor_set", 0) else 0)) * \
        (2 if kw.get("delta_win", DEFAULTS["delta_win"]) > 0 else 1)
    full = bool(kw.get("functional_set", 0))
    fun = np.full((n, (12 if full else 2) * D), np.nan)
    for i in range(n):
        x = pcm[offsets[i]:offsets[i + 1]].astype(np.float64) / 32768.0
        f = frame_lld(x, fs, **kw)
        rows.append(f)
        if len(f) and full:
            fun[i] = functionals12(f).reshape(-1)
        elif len(f):
            fun[i, :D] = f.mean(axis=0)
            fun[i, D:] = f.std(axis=0)                              # population (ddof = 0)
    return fun, rows
