"""Known-answer tests that pin the CPU oracle (oracle/) to analytic results (SURVEY.md 8c list).

The reference has no tests and no golden vectors (parity unpinned), so these are the checks that stand between
the Praat restatement and arbitrary drift.
"""
import numpy as np
import pytest

FS = 16000.0


def tone(f, dur, amp=0.3, fs=FS):
    t = np.arange(int(dur * fs)) / fs
    return amp * np.sin(2 * np.pi * f * t)


def pulse_train(f0, dur, fs=FS, amp=0.05, nharm=None):
    t = np.arange(int(dur * fs)) / fs
    nharm = nharm or int(0.45 * fs / f0)
    x = np.zeros_like(t)
    for h in range(1, nharm + 1):
        x += np.cos(2 * np.pi * h * f0 * t)
    return amp * x / nharm * 4


def test_fft_matches_numpy(orc):
    rng = np.random.default_rng(0)
    for n in (8, 1024, 4096):
        re, im = rng.normal(size=n), rng.normal(size=n)
        got = orc.fft(re, im, -1)
        np.testing.assert_allclose(got, np.fft.fft(re + 1j * im), rtol=0, atol=1e-10)
        back = orc.fft(got.real, got.imag, +1) / n
        np.testing.assert_allclose(back, re + 1j * im, atol=1e-12)


def test_frame_grid_counts_match_survey_table(orc):
    nx = 960000  # 60 s
    # AC pitch windows 3/floor, dt 5 ms
    assert orc.frame_grid(nx, FS, 3 / 60, 0.005)[0] == 11991
    assert orc.frame_grid(nx, FS, 3 / 100, 0.005)[0] == 11995
    assert orc.frame_grid(nx, FS, 3 / 75, 0.005)[0] == 11993
    assert orc.frame_grid(nx, FS, 3 / 50, 0.005)[0] == 11989
    assert orc.frame_grid(nx, FS, 3 / 30, 0.02)[0] == 2996
    assert orc.frame_grid(nx, FS, 6.4 / 50, 0.016)[0] == 3743
    assert orc.frame_grid(nx, FS, 6.4 / 60, 0.005)[0] == 11979
    assert orc.frame_grid(nx, FS, 5.5 / 60, 0.005)[0] == 11982
    # frames are centred in the file
    n, t1 = orc.frame_grid(nx, FS, 0.05, 0.005)
    assert abs((t1 + (n - 1) * 0.005 / 2) - 30.0) < 1e-9
    assert orc.frame_grid(100, FS, 0.05, 0.005) is None  # shorter than the window -> Praat throws


def test_bessel_i0(orc):
    from scipy.special import i0
    for x in (0.0, 0.5, 3.0, 3.75, 10.0, 20.2):
        assert abs(orc.bessel_i0(x) / i0(x) - 1) < 3e-7   # A&S 9.8.1/9.8.2 accuracy


def test_sinc_interpolation_band_limited(orc):
    n = 400
    k = np.arange(1, n + 1)
    y = np.sin(2 * np.pi * 0.05 * k) + 0.5 * np.cos(2 * np.pi * 0.11 * k + 0.3)
    for x in (150.25, 200.5, 217.9):
        exact = np.sin(2 * np.pi * 0.05 * x) + 0.5 * np.cos(2 * np.pi * 0.11 * x + 0.3)
        assert abs(orc.interpolate_sinc(y, x, 70) - exact) < 2e-3
        assert abs(orc.interpolate_sinc(y, x, 1) - np.interp(x, k, y)) < 1e-12
    assert orc.interpolate_sinc(y, 10.0, 70) == y[9]          # through the points
    assert orc.interpolate_sinc(y, -3.0, 70) == y[0]          # constant extrapolation
    assert orc.interpolate_sinc(y, n + 5.0, 70) == y[-1]


def test_improve_extremum_parabolic_and_sinc(orc):
    k = np.arange(1, 201)
    x0 = 100.3
    y = np.cos(2 * np.pi * 0.02 * (k - x0))
    i = int(np.argmax(y)) + 1
    v, xr = orc.improve_extremum(y, i, 1)
    assert abs(xr - x0) < 5e-3 and abs(v - 1) < 1e-4
    v, xr = orc.improve_extremum(y, i, 3)
    assert abs(xr - x0) < 1e-3 and abs(v - 1) < 1e-4
    v, xr = orc.improve_extremum(-y, i, 3, is_maximum=False)
    assert abs(xr - x0) < 1e-3 and abs(v + 1) < 1e-4


def test_quantile_and_theil(orc):
    a = np.sort(np.random.default_rng(1).normal(size=101))
    assert orc.quantile(a, 0.5) == pytest.approx(np.median(a))
    assert orc.quantile(a, 0.0) == a[0] and orc.quantile(a, 1.0) == a[-1]
    x = np.linspace(150, 4950, 49)
    y = -0.005 * x + 40 + np.random.default_rng(2).normal(scale=0.01, size=49)
    y[7] += 30.0                                               # one outlier must not move the robust fit
    m, b = orc.theil(x, y)
    assert abs(m + 0.005) < 2e-5
    m2, _ = orc.theil(x, -0.005 * x + 3.0, complete=True)
    assert abs(m2 + 0.005) < 1e-12


def test_burg_and_roots_recover_ar2_pole(orc):
    fs, f, bw = 10000.0, 700.0, 90.0
    r = np.exp(-np.pi * bw / fs)
    a1, a2 = 2 * r * np.cos(2 * np.pi * f / fs), -r * r
    rng = np.random.default_rng(3)
    e = rng.normal(size=20000)
    x = np.zeros_like(e)
    for i in range(2, len(e)):
        x[i] = a1 * x[i - 1] + a2 * x[i - 2] + e[i]
    a, xms = orc.burg(x[1000:], 2)
    assert abs(a[0] - a1) < 0.01 and abs(a[1] - a2) < 0.01 and xms > 0
    z = orc.roots([-a[1], -a[0], 1.0])
    zf = np.abs(np.angle(z)) * fs / (2 * np.pi)
    zb = -np.log(np.abs(z) ** 2) * fs / (2 * np.pi)
    assert np.all(np.abs(zf - f) < 5) and np.all(np.abs(zb - bw) < 15)
    # generic degree-10 polynomial vs numpy
    c = rng.normal(size=11)
    c[-1] = 1.0
    got = np.sort_complex(orc.roots(c))
    want = np.sort_complex(np.roots(c[::-1]))
    np.testing.assert_allclose(got, want, atol=1e-8)


def test_intensity_of_sine_is_rms_level(orc):
    A = 0.2
    x = tone(440.0, 1.0, amp=A)
    c, x1 = orc.intensity(x, FS, 100.0, 0.005)
    want = 20 * np.log10(A / np.sqrt(2) / 2e-5)
    mid = c[len(c) // 4: 3 * len(c) // 4]
    assert np.max(np.abs(mid - want)) < 1e-3
    st = orc.intensity_stats(x, FS, 100.0, 0.005)
    assert abs(st[3] - want) < 1e-2 and st[0] <= st[2] <= st[1]


def test_pitch_of_pulse_train_ac_and_cc(orc):
    for f0 in (110.0, 220.0):
        x = pulse_train(f0, 1.0)
        p = orc.pitch(x, FS, method=0, dt=0.005, floor=75.0, ceiling=600.0)
        assert np.all(p["freq"][5:-5] > 0)
        assert np.max(np.abs(p["freq"][5:-5] - f0)) < 0.05
        assert np.min(p["strength"][5:-5]) > 0.95
        q = orc.pitch(x, FS, method=2, dt=0.005, floor=75.0, ppw=1.0, ceiling=600.0)
        assert np.max(np.abs(q["freq"][5:-5] - f0)) < 0.05
        h = orc.hnr(x, FS, 0.005, 75.0, 0.1, 4.5)
        assert h > 30.0                                        # nearly perfectly periodic


def test_octave_cost_prefers_fundamental(orc):
    # fundamental + strong second harmonic: candidates at T0 and T0/2 ... AC must keep f0, not 2*f0
    t = np.arange(int(FS)) / FS
    x = 0.1 * np.sin(2 * np.pi * 120 * t) + 0.2 * np.sin(2 * np.pi * 240 * t + 0.4)
    p = orc.pitch(x, FS, method=0, dt=0.01, floor=75.0, ceiling=600.0)
    assert np.max(np.abs(p["freq"][3:-3] - 120.0)) < 0.5


def test_silent_and_constant_clips_take_the_reference_fallbacks(orc):
    z = np.zeros(int(2 * FS))
    assert orc.pitch_values(z) == (75.0, 500.0, True)          # :146 no voiced frames
    out, st = orc.extract_f64(z)
    assert np.isnan(out[5]) and np.isnan(out[6])               # mean_F0 / sd undefined
    assert np.isnan(out[12]) and np.all(np.isnan(out[13:21])) and np.all(np.isnan(out[21:25]))
    assert out[7] == pytest.approx(-300.0)                     # intensity floor
    out2, _ = orc.extract_f64(np.full(int(2 * FS), 0.25))
    assert np.isnan(out2[5])
    # too short for every analysis window: all groups NaN, no crash
    out3, _ = orc.extract_f64(np.zeros(100))
    assert np.all(np.isnan(out3))


def test_gated_tone_pauses_on_16ms_grid(orc):
    rng = np.random.default_rng(5)
    seg = []
    for k in range(4):
        seg.append(tone(150.0, 1.0, amp=0.3) * (1 + 0.5 * np.sin(2 * np.pi * 4 * np.arange(int(FS)) / FS)))
        seg.append(rng.normal(scale=1e-4, size=int(0.5 * FS)))
    x = np.concatenate(seg)
    out, ok = orc.speechrate(x)
    assert ok
    # 4 sounding stretches -> 3 pauses; pause durations ~0.5 s each, quantised to the 16 ms intensity grid
    assert 0.38 < out[4] <= 0.5                                # the 128 ms Kaiser window smears the edges
    q = out[4] * 3 / 0.016
    assert abs(q - round(q)) < 1e-6
    assert 0.6 < out[2] < 0.8
    assert out[0] > 0 and out[1] >= out[0]


def test_flat_harmonic_spectrum_moments(orc):
    x = pulse_train(125.0, 1.5, nharm=60)                       # equal-amplitude harmonics up to 7.5 kHz
    m = orc.moments(x, FS, 60.0, 250.0)
    assert abs(m[0] - 2500.0) < 60.0
    assert abs(m[1] - 5000.0 / np.sqrt(12)) < 40.0
    assert abs(m[2]) < 0.05 and abs(m[3] + 1.2) < 0.05


def test_cpps_of_pulse_train_is_prominent(orc):
    x = pulse_train(125.0, 0.6, nharm=30)
    v = orc.cpps_segment(x, FS)
    assert v is not None and v > 15.0
    noise = np.random.default_rng(7).normal(scale=0.05, size=int(0.6 * FS))
    vn = orc.cpps_segment(noise, FS)
    assert vn < v - 8.0


def test_resample_preserves_in_band_sine(orc):
    x = tone(1000.0, 0.5, amp=0.25)
    y, x1 = orc.resample(x, FS, 10000.0, 50)
    assert len(y) == 5000
    t = x1 + np.arange(len(y)) / 10000.0
    want = 0.25 * np.sin(2 * np.pi * 1000.0 * (t - 0.5 / FS))    # Sound sample k sits at (k + 0.5)/fs
    assert np.max(np.abs(y[300:-300] - want[300:-300])) < 2e-4
    z, _ = orc.resample(tone(7000.0, 0.5), FS, 10000.0, 50)     # above the new Nyquist: removed
    assert np.max(np.abs(z[300:-300])) < 2e-3


def test_formants_of_two_resonator_signal(orc):
    fs = 16000.0
    rng = np.random.default_rng(11)
    n = int(1.0 * fs)
    src = np.zeros(n)
    src[::128] = 1.0                                            # 125 Hz pulse train
    src += rng.normal(scale=0.01, size=n)
    y = src
    for f, bw in ((600.0, 80.0), (1700.0, 110.0)):
        r = np.exp(-np.pi * bw / fs)
        a1, a2 = 2 * r * np.cos(2 * np.pi * f / fs), -r * r
        out = np.zeros(n)
        for i in range(2, n):
            out[i] = a1 * out[i - 1] + a2 * out[i - 2] + y[i]
        y = out
    y = 0.2 * y / np.max(np.abs(y))
    fm = orc.formants(y, fs)
    mid = slice(20, -20)
    f = fm["f"][mid]
    bw = fm["bw"][mid]
    # with 10 poles on a 2-resonance signal the spare poles are broad; the two sharp ones are the resonators
    sharp = np.where(bw < 300.0, f, np.nan)
    f1 = np.nanmin(sharp, axis=1)
    f2 = np.nanmax(np.where(sharp < 2500.0, sharp, np.nan), axis=1)
    assert abs(np.nanmedian(f1) - 600.0) < 40.0
    assert abs(np.nanmedian(f2) - 1700.0) < 60.0


def test_exact_doubling_takes_the_upsample_route(orc):
    """Sound_resample hands upfactor == 2 to Sound_upsample: twice the samples, dx / 2, x1 - dx / 4; the even samples are the
    original ones (away from the ends) and the odd ones the band-limited values half a sample later."""
    fs = 8000.0
    t = np.arange(4001) / fs
    x = 0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3100 * t + 0.7)
    y, x1 = orc.resample(x, fs, 16000.0, 50)
    assert len(y) == 2 * len(x) and abs(x1 - (0.5 / fs - 0.25 / fs)) < 1e-15
    tm = np.arange(len(y)) / 16000.0
    ref = 0.3 * np.sin(2 * np.pi * 440 * tm) + 0.2 * np.sin(2 * np.pi * 3100 * tm + 0.7)
    assert np.max(np.abs(y - ref)[600:-600]) < 1e-4
    # a component above 95 % of the old Nyquist frequency is attenuated by the taper
    z = np.sin(2 * np.pi * 3950 * t)
    yz, _ = orc.resample(z, fs, 16000.0, 50)
    assert 0.1 < np.sqrt(2 * np.mean(yz[2000:-2000] ** 2)) < 0.6


def test_harmonicity_equals_signal_to_noise_ratio(orc):
    """A periodic signal plus white noise has autocorrelation r(T0) = S / (S + N) at its period, so the harmonics-to-noise
    ratio 10 log10(r / (1 - r)) that Sound_to_Harmonicity_cc reports is the SNR in dB (Boersma 1993, eq. 4)."""
    rng = np.random.default_rng(12)
    x = pulse_train(125.0, 2.0, amp=0.2)
    ps = np.mean(x ** 2)
    for snr_db in (10.0, 20.0):
        noise = rng.normal(size=len(x))
        noise *= np.sqrt(ps / 10 ** (snr_db / 10) / np.mean(noise ** 2))
        h = orc.hnr(x + noise, FS, 0.005, 75.0, 0.1, 4.5)
        assert abs(h - snr_db) < 1.5, (snr_db, h)


def test_glottal_pulses_sit_one_period_apart(orc):
    """Sound_Pitch_to_PointProcess_cc on a strictly periodic source: consecutive pulses are exactly one period apart and their
    number is duration * f0 (to within the voiced-stretch ends)."""
    f0 = 160.0
    x = pulse_train(f0, 1.5, amp=0.2)
    for method, ppw in ((0, 3.0), (2, 1.0)):
        t = orc.pulses(x, FS, method, 0.005, 75.0, ppw, 0.45, 600.0)
        assert abs(len(t) - 1.5 * f0) <= 8
        d = np.diff(t)
        assert np.max(np.abs(d - 1.0 / f0)) < 2e-5
        assert np.all(np.diff(t) > 0)


def test_ltas_slope_of_a_two_band_signal(orc):
    """'Get slope 50 1000 1000 4000 dB' = mean level of the high band minus that of the low band.  A periodic source whose
    harmonics above 1 kHz are 20 dB weaker than those below gives about -20 dB; equal levels give about 0 dB."""
    f0 = 125.0
    t = np.arange(int(2.0 * FS)) / FS
    for att_db in (0.0, 20.0):
        x = np.zeros_like(t)
        for hnum in range(1, 33):                               # harmonics up to 4 kHz
            a = 1.0 if hnum * f0 < 1000.0 else 10 ** (-att_db / 20)
            x += a * np.cos(2 * np.pi * hnum * f0 * t)
        x *= 0.02
        bands, out2 = orc.ltas(x, FS, 75.0, 600.0)
        assert abs(out2[0] - (-att_db)) < 2.5, (att_db, out2)
        assert out2[1] <= 0.0005 and (att_db == 0.0 or out2[1] < 0.0)          # robust tilt in dB/Hz: falling spectrum


def test_pitch_statistics_of_a_two_level_contour(orc):
    """'Get mean' / 'Get standard deviation (semitones)' over voiced frames: half a second at 100 Hz, half a second one octave
    higher -> mean 150 Hz, sample standard deviation of 12 log2(f) close to 6 semitones."""
    x = np.concatenate([pulse_train(100.0, 0.6, amp=0.2), pulse_train(200.0, 0.6, amp=0.2)])
    f, st = orc.extract_f64(x, FS)
    assert abs(f[5] - 150.0) < 3.0
    assert abs(f[6] - 6.0) < 0.25
