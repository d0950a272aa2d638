"""Seeded synthetic 16 kHz voiced-speech generator.

The Androids corpus the reference ran on (notebooks/01_feature_extraction_setup.ipynb:51,56) is not available
offline, so benchmarks and parity tests use this generator: phrases and pauses, syllable-rate amplitude
modulation with 6-10 dB dips, a harmonic glottal source (-12 dB/oct) with per-period jitter and shimmer, four
slowly varying formant resonances, ~15 % noise-excited (unvoiced) syllables, additive noise, peak normalisation
and int16 quantisation.  Everything the 25 MSHDS columns look at (pauses, syllable nuclei, voicing, F0 range of
both speaker classes, formants, spectral tilt) is exercised.

Schedules and per-period perturbations come from a per-clip numpy Generator (device independent); per-sample
noise is a counter-based integer hash evaluated with torch ops on the target device.
"""
from __future__ import annotations

import math

import numpy as np
import torch

FS = 16000
BASE_SEED = 20251018


def _hash_uniform(idx: torch.Tensor, seed: int) -> torch.Tensor:
    """Counter-based uniform(0,1) from int64 sample indices (integer arithmetic only: device independent)."""
    m = 0xFFFFFFFF
    x = (idx * 747796405 + (seed & m) * 2891336453 + 12345) & m
    x = ((x >> 16) ^ x) * 0x45D9F3B & m
    x = ((x >> 16) ^ x) * 0x45D9F3B & m
    x = (x >> 16) ^ x
    return (x.to(torch.float64) + 0.5) / 4294967296.0


def _gauss(idx: torch.Tensor, seed: int) -> torch.Tensor:
    u = _hash_uniform(idx, seed) + _hash_uniform(idx, seed + 7919) + _hash_uniform(idx, seed + 104729) + \
        _hash_uniform(idx, seed + 1299709)
    return (u - 2.0) * math.sqrt(3.0)


def _schedule(rng: np.random.Generator, duration: float, style: str = "plain"):
    """Phrase / pause / syllable timeline -> arrays of syllable (start, end, voiced, dip_dB, formants).

    style "plain": phrases of 0.8-3 s separated by pauses of 0.35-0.9 s (all longer than the 0.3 s minimum pause of
    _speechrate).  style "stress": short bursts (0.12-1.5 s) and gaps of 0.03-0.6 s, 30 % unvoiced syllables -- sounding /
    silent intervals below the 0.1 s / 0.3 s minima that must be cut and merged, voiced runs whose +-50 ms extension overlaps
    the next one, many voiced/unvoiced transitions for the path finder."""
    stress = style == "stress"
    syl = []
    t = float(rng.uniform(0.15, 0.5))          # leading pause
    while t < duration - 0.2:
        phrase_end = min(duration - 0.15, t + float(rng.uniform(0.12, 1.5) if stress else rng.uniform(0.8, 3.0)))
        rate = float(rng.uniform(3.5, 5.5))
        f_form = np.array([500.0, 1500.0, 2500.0, 3500.0]) + rng.normal(0, 1, 4) * np.array([80, 150, 150, 150.0])
        while t < phrase_end - 0.08:
            d = min(phrase_end - t, float(rng.uniform(0.8, 1.2)) / rate)
            f_form = f_form + rng.normal(0, 1, 4) * np.array([40, 80, 60, 60.0])
            f_form = np.clip(f_form, [300, 1000, 2100, 3100], [800, 2000, 2900, 3900])
            syl.append((t, t + d, rng.random() > (0.30 if stress else 0.15), float(rng.uniform(6.0, 10.0)), f_form.copy()))
            t += d
        t = phrase_end + float(rng.uniform(0.03, 0.6) if stress else rng.uniform(0.35, 0.9))
    return syl


def synth_clip(index: int, duration: float, device: str | torch.device = "cpu", fs: int = FS, style: str = "plain") -> torch.Tensor:
    """One clip as an int16 tensor of round(duration*fs) samples on `device` (style: see _schedule)."""
    rng = np.random.default_rng(BASE_SEED + index)
    n = int(round(duration * fs))
    dev = torch.device(device)
    base_f0 = float(rng.uniform(95, 135)) if index % 2 == 0 else float(rng.uniform(180, 230))
    syl = _schedule(rng, duration, style)
    if not syl:
        syl = [(0.1 * duration, 0.9 * duration, True, 8.0, np.array([500.0, 1500.0, 2500.0, 3500.0]))]
    starts = torch.tensor([s[0] for s in syl], dtype=torch.float64, device=dev)
    ends = torch.tensor([s[1] for s in syl], dtype=torch.float64, device=dev)
    voiced = torch.tensor([1.0 if s[2] else 0.0 for s in syl], dtype=torch.float64, device=dev)
    dips = torch.tensor([10 ** (-s[3] / 20) for s in syl], dtype=torch.float64, device=dev)
    forms = torch.tensor(np.stack([s[4] for s in syl]), dtype=torch.float64, device=dev)   # [S,4]

    idx = torch.arange(n, device=dev, dtype=torch.int64)
    t = idx.to(torch.float64) / fs
    si = torch.searchsorted(ends, t, right=False).clamp_(max=len(syl) - 1)
    inside = (t >= starts[si]) & (t < ends[si])
    u = ((t - starts[si]) / (ends[si] - starts[si])).clamp_(0, 1)
    env = (dips[si] + (1 - dips[si]) * 0.5 * (1 - torch.cos(2 * math.pi * u))) * inside
    # soften the phrase edges (first / last 30 ms of a run of syllables)
    v = voiced[si] * inside

    # F0 contour: vibrato-like slow modulation + random walk (per 50 ms knots, linearly interpolated)
    nk = int(duration / 0.05) + 3
    walk = np.cumsum(rng.normal(0, 0.5 * math.sqrt(0.05), nk))
    walk = torch.tensor(walk - walk.mean(), dtype=torch.float64, device=dev)
    kpos = t / 0.05
    k0 = kpos.floor().long().clamp_(max=nk - 2)
    w = kpos - k0
    semis = 2.0 * torch.sin(2 * math.pi * 0.3 * t + float(rng.uniform(0, 2 * math.pi))) + walk[k0] * (1 - w) + walk[k0 + 1] * w
    f0 = base_f0 * torch.pow(2.0, semis / 12.0)

    # per-period jitter / shimmer: sample-and-hold noise indexed by the running period count
    cyc0 = torch.cumsum(f0 / fs, 0)
    nper = int(cyc0[-1].item()) + 4
    jit = torch.tensor(1.0 + 0.005 * rng.normal(0, 1, nper), dtype=torch.float64, device=dev)
    shim = torch.tensor(1.0 + 0.03 * rng.normal(0, 1, nper), dtype=torch.float64, device=dev)
    pidx = cyc0.floor().long().clamp_(max=nper - 1)
    f0j = f0 * jit[pidx]
    cyc = torch.cumsum(f0j / fs, 0)
    amp_p = shim[cyc.floor().long().clamp_(max=nper - 1)]

    F = forms[si]                                              # [n,4]
    bw = torch.tensor([80.0, 100.0, 120.0, 150.0], dtype=torch.float64, device=dev)
    r = torch.exp(-math.pi * bw / fs)                          # [4]
    th = 2 * math.pi * F / fs                                  # [n,4]
    gain0 = 1 - 2 * r * torch.cos(th) + r * r                  # DC-normalised resonators

    def tract(freq):                                           # |H(f)| of the 4-resonator cascade
        om = (2 * math.pi * freq / fs).unsqueeze(1)
        a = torch.sqrt(1 - 2 * r * torch.cos(om - th) + r * r)
        b = torch.sqrt(1 - 2 * r * torch.cos(om + th) + r * r)
        return (gain0 / (a * b)).prod(dim=1)

    voiced_sig = torch.zeros(n, dtype=torch.float64, device=dev)
    hmax = int(7000.0 / (base_f0 * 0.7))
    ph = 2 * math.pi * cyc
    for h in range(1, hmax + 1):
        fh = f0j * h
        a_h = tract(fh) / (h * h) * (fh < 7000.0)
        voiced_sig += a_h * torch.sin(h * ph + 0.7 * h)
    voiced_sig *= amp_p

    seed = BASE_SEED + 7 * index
    white = _gauss(idx, seed)
    # unvoiced syllables: noise through a gentle high-frequency emphasis (first difference mix)
    frict = white - 0.6 * torch.roll(white, 1)
    vs = voiced_sig * v * env
    us = frict * (inside & (v == 0)) * env
    vrms = torch.sqrt((vs * vs).sum() / v.sum().clamp_(min=1.0) + 1e-30)
    us = us * (0.35 * vrms / 1.2)
    sig = vs + us
    # additive noise: 25 dB SNR inside syllables, -55 dBFS floor everywhere (never exactly silent)
    snr_noise = vrms * 10 ** (-25 / 20)
    sig = sig + _gauss(idx, seed + 31) * (snr_noise * inside + 0.0)
    peak = sig.abs().max().clamp_(min=1e-9)
    sig = sig * (0.5 / peak)
    sig = sig + _gauss(idx, seed + 57) * (10 ** (-55 / 20))
    pcm = torch.round(sig * 32767.0).clamp_(-32768, 32767).to(torch.int16)
    return pcm


def synth_batch(n_clips: int, duration, device: str | torch.device = "cpu", start_index: int = 0, unique: int | None = None):
    """Packed batch: (pcm int16 [total], offsets int64 [n_clips+1]).

    `duration` is a float (all clips equal) or a sequence of per-clip durations.  `unique` limits the number of
    distinct clips that are synthesised (the rest are cyclic repeats) to bound set-up time for very large batches.
    """
    durs = [float(duration)] * n_clips if np.isscalar(duration) else [float(d) for d in duration]
    assert len(durs) == n_clips
    cache = {}
    clips = []
    for i in range(n_clips):
        key = (i % unique if unique else i, durs[i])
        if key not in cache:
            cache[key] = synth_clip(start_index + key[0], durs[i], device)
        clips.append(cache[key])
    lens = torch.tensor([0] + [c.numel() for c in clips], dtype=torch.int64)
    offsets = torch.cumsum(lens, 0)
    pcm = torch.cat(clips) if clips else torch.zeros(0, dtype=torch.int16, device=device)
    return pcm, offsets
