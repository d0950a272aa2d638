"""CPU-side tests (no GPU): oracle vs frozen golden vectors, C-ABI export check, host logic of the drop-in, sharding."""
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "mshds_golden_v1.npz")


def test_oracle_reproduces_golden(orc):
    g = np.load(GOLDEN)
    feats, status = orc.extract(g["pcm"], g["offsets"], 16000.0, nthreads=os.cpu_count() or 1)
    assert list(g["feature_names"]) == orc.FEATURE_NAMES
    assert np.array_equal(np.isnan(feats), np.isnan(g["features"]))
    np.testing.assert_allclose(feats, g["features"], rtol=1e-9, atol=1e-12, equal_nan=True)
    assert np.array_equal(status, g["status"])


def test_golden_values_are_plausible_against_reference_printout():
    # the only numbers the reference repo shows (5 Androids files, inputs unavailable): unit sanity.
    g = np.load(GOLDEN)
    f = g["features"][:4]
    names = list(g["feature_names"])
    col = lambda n: f[:, names.index(n)]
    assert np.all((col("Speaking_Rate") > 0.5) & (col("Speaking_Rate") < 8))
    assert np.all((col("Phonation_Ratio") > 0.3) & (col("Phonation_Ratio") <= 1.0))
    assert np.all((col("mean_F0") > 60) & (col("mean_F0") < 400))
    assert np.all((col("mean_dB") > 40) & (col("mean_dB") < 95))          # dB re 2e-5 Pa
    assert np.all((col("Spectral_Tilt") < 0) & (col("Spectral_Tilt") > -0.05))   # dB/Hz
    assert np.all((col("Cepstral_Peak_Prominence") > 4) & (col("Cepstral_Peak_Prominence") < 40))
    assert np.all((col("mean_F1_Loc") > 200) & (col("mean_F1_Loc") < 1000))
    # pause durations live on the 16 ms intensity grid (App. B note)
    mp, pr = col("Mean_Pause_Duration"), col("Pause_Rate")
    assert np.all(mp >= 0)


def test_c_abi_library_exports_every_declared_symbol():
    from robust_speech_analysis_framework_b200 import _lib
    header = open(os.path.join(ROOT, "include", "mshds_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mshds_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    lib = _lib.load()
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/mshds_b200.h but not exported"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from robust_speech_analysis_framework_b200 import _lib
    with pytest.raises(_lib.MshdsError):
        _lib.Extractor(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "robust_speech_analysis_framework_b200")
    pat = re.compile(r"^\s*(import\s+oracle|from\s+oracle|from\s+\.\.?oracle)|libmshds_oracle|#include\s+\".*oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dirpath, fn)).read()), fn


def _write_wav(path, x, fs=16000, nch=1):
    import wave
    with wave.open(path, "wb") as w:
        w.setnchannels(nch)
        w.setsampwidth(2)
        w.setframerate(fs)
        w.writeframes(np.asarray(x, dtype="<i2").tobytes())


def test_wav_reader(tmp_path):
    from robust_speech_analysis_framework_b200.mshds_extractor import AudioLoadError, read_wav_mono_int16
    x = (np.arange(-500, 500) * 13).astype(np.int16)
    p = str(tmp_path / "a.wav")
    _write_wav(p, x)
    y, fs = read_wav_mono_int16(p)
    assert fs == 16000 and np.array_equal(x, y)
    st = np.stack([x, -x], axis=1).ravel()
    _write_wav(p, st, nch=2)
    y, _ = read_wav_mono_int16(p)
    assert np.all(y == 0)
    with pytest.raises(AudioLoadError):
        read_wav_mono_int16(str(tmp_path / "missing.wav"))
    # the reader of the drop-in: mono 16-bit stays int16, anything else becomes float64 exactly as Praat holds it
    import wave
    from robust_speech_analysis_framework_b200.mshds_extractor import read_wav_mono
    _write_wav(p, x)
    y, _ = read_wav_mono(p)
    assert y.dtype == np.int16 and np.array_equal(x, y)
    st2 = np.stack([x, x + 1], axis=1).ravel()                 # channel mean = x + 0.5: not an int16 value
    _write_wav(p, st2, nch=2)
    y, _ = read_wav_mono(p)
    assert y.dtype == np.float64 and np.array_equal(y, (x.astype(np.float64) + 0.5) / 32768.0)
    v24 = np.array([-8388608, -1, 0, 1, 8388607, 123456], dtype=np.int64)
    with wave.open(p, "wb") as w:
        w.setnchannels(1); w.setsampwidth(3); w.setframerate(22050)
        w.writeframes(b"".join(int(v & 0xffffff).to_bytes(3, "little") for v in v24))
    y, fs = read_wav_mono(p)
    assert fs == 22050 and y.dtype == np.float64 and np.array_equal(y, v24 / 8388608.0)


class _FakeExtractor:
    """Stands in for the CUDA handle so the DataFrame / error-convention logic can be tested without a GPU."""
    def __init__(self):
        self.calls = []

    def extract_host(self, pcm, offsets, sample_rate=16000):
        n = len(offsets) - 1
        self.calls.append((int(sample_rate), n))
        if sample_rate == 8000:
            raise RuntimeError("simulated device failure")             # a failing device call must give NaN rows + messages
        out = np.zeros((n, 25))
        for i in range(n):
            seg = pcm[offsets[i]:offsets[i + 1]].astype(np.float64)
            out[i] = seg.sum() + np.arange(25)
        return out, np.zeros(n, np.uint32)

    def extract_host_f64(self, samples, offsets, sample_rate=16000):
        out, st = self.extract_host(np.asarray(samples) * 32768.0, offsets, sample_rate)
        self.calls[-1] = (self.calls[-1][0], self.calls[-1][1], "f64")
        return out, st


def test_dataframe_contract_matches_reference(tmp_path, monkeypatch, capsys):
    import pandas as pd
    from robust_speech_analysis_framework_b200 import mshds_extractor as mx
    fake = _FakeExtractor()
    monkeypatch.setattr(mx, "get_extractor", lambda device=0: fake)
    a, b = np.full(4000, 3, np.int16), np.full(2000, -5, np.int16)
    pa, pb = str(tmp_path / "s1" / "a.wav"), str(tmp_path / "b.wav")
    os.makedirs(os.path.dirname(pa))
    _write_wav(pa, a)
    _write_wav(pb, b)
    _write_wav(str(tmp_path / "c44.wav"), a, fs=44100)
    _write_wav(str(tmp_path / "d8.wav"), b, fs=8000)
    df = pd.DataFrame({"other": [1, 2, 3, 4, 5], "filepath": [pa, str(tmp_path / "nope.wav"), pb, str(tmp_path / "c44.wav"),
                                                             str(tmp_path / "d8.wav")]})
    out = mx.extract_mshds_features(df, verbose=True)
    # mshds_extractor.py:240 smoke check shape (n, 26); column order :397-404; filename = basename (:410)
    assert list(out.columns) == ["filename"] + mx.FEATURE_NAMES and out.shape == (5, 26)
    assert list(out["filename"]) == ["a.wav", "nope.wav", "b.wav", "c44.wav", "d8.wav"]
    assert out.iloc[0]["Speaking_Rate"] == 12000.0 and out.iloc[2]["Spectral_Kurtosis"] == -10000.0 + 24
    # one device call per sampling frequency, each told its rate (the library does resample(16000, 50), :418-419)
    # (the failing 8 kHz call is retried once, recording by recording)
    assert sorted(fake.calls) == [(8000, 1), (8000, 1), (16000, 2), (44100, 1)]
    assert out.iloc[3]["Speaking_Rate"] == 12000.0
    assert out.iloc[1][mx.FEATURE_NAMES].isna().all() and out.iloc[4][mx.FEATURE_NAMES].isna().all()
    printed = capsys.readouterr().out
    assert "ERROR processing file 'nope.wav'" in printed and "ERROR processing file 'd8.wav'" in printed     # :452-453
    quiet = mx.extract_mshds_features(df, verbose=False)
    assert capsys.readouterr().out == "" and quiet.shape == (5, 26)
    assert out[mx.FEATURE_NAMES].dtypes.map(lambda d: d == np.float64).all()
    # custom column name (:379 audio_file_column)
    out2 = mx.extract_mshds_features(df.rename(columns={"filepath": "p"}), audio_file_column="p", verbose=False)
    assert out2.equals(quiet)
    # drop-in import path of the notebooks (01_feature_extraction_setup.ipynb:31)
    sys.path.insert(0, ROOT)
    from src.mshds_extractor import extract_mshds_features as shim
    assert shim is mx.extract_mshds_features


def test_a_failing_device_call_only_costs_the_recording_that_causes_it(tmp_path, monkeypatch, capsys):
    """Reference convention (:450-457): a file that cannot be processed gives ONE NaN row.  A batched device call that fails
    is retried recording by recording, device errors are surfaced with warnings.warn even when verbose is off."""
    import pandas as pd
    from robust_speech_analysis_framework_b200 import mshds_extractor as mx

    class Poisoned(_FakeExtractor):
        def extract_host(self, pcm, offsets, sample_rate=16000):
            if (np.asarray(pcm) == 666).any():
                self.calls.append((int(sample_rate), len(offsets) - 1))
                raise RuntimeError("simulated out of memory")
            return super().extract_host(pcm, offsets, sample_rate)

    fake = Poisoned()
    monkeypatch.setattr(mx, "get_extractor", lambda device=0: fake)
    paths = []
    for k, v in enumerate((3, 666, 5)):
        p = str(tmp_path / f"f{k}.wav")
        _write_wav(p, np.full(1000, v, np.int16))
        paths.append(p)
    with pytest.warns(RuntimeWarning):
        out = mx.extract_mshds_features(pd.DataFrame({"filepath": paths}), verbose=False)
    assert capsys.readouterr().out == ""
    assert out.iloc[0]["Speaking_Rate"] == 3000.0 and out.iloc[2]["Speaking_Rate"] == 5000.0
    assert out.iloc[1][mx.FEATURE_NAMES].isna().all()
    # several devices: LPT split, one thread + handle per device, rows back in input order
    fake2 = _FakeExtractor()
    seen = []
    monkeypatch.setattr(mx, "get_extractor", lambda device=0: (seen.append(device), fake2)[1])
    out2 = mx.extract_mshds_features(pd.DataFrame({"filepath": [paths[0], paths[2], paths[0]]}), verbose=False, devices=[0, 1])
    assert list(out2["Speaking_Rate"]) == [3000.0, 5000.0, 3000.0] and {0, 1} <= set(seen)


def test_float_wav_and_aiff_are_decoded_and_other_containers_are_a_distinct_error(tmp_path):
    from scipy.io import wavfile
    from robust_speech_analysis_framework_b200.mshds_extractor import AudioLoadError, read_wav_mono
    x = (0.25 * np.sin(np.arange(800) * 0.1)).astype(np.float32)
    p = str(tmp_path / "f32.wav")
    wavfile.write(p, 16000, x)                                   # IEEE-float WAV: stdlib wave refuses it
    y, fs = read_wav_mono(p)
    assert fs == 16000 and y.dtype == np.float64 and np.array_equal(y, x.astype(np.float64))
    st = np.stack([x, -x], axis=1)
    wavfile.write(p, 22050, st)
    y, fs = read_wav_mono(p)
    assert fs == 22050 and np.allclose(y, 0.0)                   # convert_to_mono averages the channels (:416-417)
    q = str(tmp_path / "x.flac")
    open(q, "wb").write(b"fLaC" + bytes(64))
    with pytest.raises(AudioLoadError, match="FLAC"):
        read_wav_mono(q)
    # with the optional `soundfile` package present, other containers are decoded through it (stereo -> channel mean)
    import sys
    import types
    fake = types.ModuleType("soundfile")
    fake.read = lambda path, dtype="float64", always_2d=True: (np.array([[0.5, -0.5], [0.25, 0.75], [-1.0, 0.0]]), 22050)
    sys.modules["soundfile"] = fake
    try:
        y, fs = read_wav_mono(q)
    finally:
        del sys.modules["soundfile"]
    assert fs == 22050 and y.dtype == np.float64 and list(y) == [0.0, 0.5, -0.5]


def test_lpt_sharding_is_a_partition_and_balanced():
    from robust_speech_analysis_framework_b200.sharding import lpt_assign, pack_subset
    rng = np.random.default_rng(0)
    lens = rng.integers(60, 600, size=230) * 16000
    for w in (1, 2, 4, 8):
        parts = lpt_assign(lens, w)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(230))
        loads = [int(lens[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= lens.max()
    pcm = np.arange(30, dtype=np.int16)
    off = np.array([0, 10, 10, 25, 30])
    d, o = pack_subset(pcm, off, [3, 0, 1])
    assert list(o) == [0, 5, 15, 15] and list(d[:5]) == list(range(25, 30))
    # equal-length clips are interchangeable: every rank gets ONE run of consecutive clips, which packs as a view (no copy)
    eq = np.full(1000, 32000)
    parts = lpt_assign(eq, 3)
    assert [len(p) for p in parts] == [334, 333, 333]
    assert all(p == list(range(p[0], p[0] + len(p))) for p in parts) and parts[1][0] == 334
    big = np.arange(64, dtype=np.int16)
    offs = np.arange(0, 65, 8)
    v, o = pack_subset(big, offs, [2, 3, 4])
    assert np.shares_memory(v, big) and list(v) == list(range(16, 40)) and list(o) == [0, 8, 16, 24]
    c, o = pack_subset(big, offs, [0, 1, 5, 6, 3])                       # runs (0,1), (5,6), (3): copied run by run, order kept
    assert not np.shares_memory(c, big) and list(o) == [0, 8, 16, 24, 32, 40]
    assert list(c) == list(range(0, 16)) + list(range(40, 56)) + list(range(24, 32))
    e, o = pack_subset(big, offs, [])
    assert len(e) == 0 and list(o) == [0]
    # mixed lengths: the regrouping of equal lengths never changes a rank's load
    mixed = np.array([5, 9, 5, 5, 9, 2, 5, 9, 2, 5]) * 1000
    parts = lpt_assign(mixed, 3)
    assert sorted(i for p in parts for i in p) == list(range(10))
    assert sorted(int(mixed[p].sum()) for p in parts) == [18000, 19000, 19000]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, pcm, off, q):
    import torch.distributed as dist
    from robust_speech_analysis_framework_b200.sharding import extract_sharded
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    out, st = extract_sharded(pcm, off, _FakeExtractor().extract_host, rank, world)
    if rank == 0:
        q.put((out, st))
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_gloo_shard_and_gather_equals_single_process():
    import torch.multiprocessing as mp
    from robust_speech_analysis_framework_b200.sharding import extract_sharded
    rng = np.random.default_rng(1)
    lens = rng.integers(100, 3000, size=13)
    lens[4] = 0                                      # an empty recording must keep its row
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pcm = rng.integers(-2000, 2000, size=int(off[-1])).astype(np.int16)
    want, wst = extract_sharded(pcm, off, _FakeExtractor().extract_host, 0, 1)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, pcm, off, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, gst = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert np.array_equal(got, want) and np.array_equal(gst, wst)


def test_synth_is_deterministic_and_well_formed():
    from robust_speech_analysis_framework_b200.synth import synth_batch, synth_clip
    a, b = synth_clip(3, 1.5), synth_clip(3, 1.5)
    assert a.dtype.is_signed and a.numel() == 24000 and bool((a == b).all())
    assert int(a.abs().max()) > 8000 and int(a.abs().max()) <= 32767
    pcm, off = synth_batch(3, [0.5, 0.25, 0.75])
    assert list(off) == [0, 8000, 12000, 24000] and pcm.numel() == 24000


# ------------------------------------------------------------------------------------------------ session aggregation (8f-3)
AGG_GOLDEN = os.path.join(ROOT, "tests", "golden", "session_agg_golden_v1.npz")


def _agg_golden_frames():
    import pandas as pd
    g = np.load(AGG_GOLDEN, allow_pickle=True)
    clip = pd.DataFrame(g["clip_values"], columns=list(g["clip_columns"]))
    clip.insert(0, "filename", list(g["clip_filenames"]))
    meta = pd.DataFrame({"filename": list(g["meta_filenames"]), "unique_participant_id": list(g["meta_ids"])})
    return g, clip, meta


def test_session_agg_restatement_is_bit_identical_to_the_reference_output():
    """oracle/session_agg.py against vectors produced by the reference's own aggregate_clip_features (src/utils.py:7-58)."""
    from oracle import session_agg
    g, clip, meta = _agg_golden_frames()
    out = session_agg.aggregate_clip_features(clip, meta)
    assert list(out["unique_participant_id"]) == list(g["out_ids"])
    assert list(out.columns[1:]) == list(g["out_columns"])
    got = out.iloc[:, 1:].to_numpy(dtype=np.float64)
    assert np.array_equal(np.isnan(got), np.isnan(g["out_values"]))
    assert np.array_equal(got, g["out_values"], equal_nan=True)           # bit for bit
    assert np.isnan(g["out_values"]).any() and out.shape == (23, 51)


def test_session_agg_restatement_matches_live_pandas_on_random_groups():
    import pandas as pd
    from oracle import session_agg
    rng = np.random.default_rng(5)
    for trial in range(4):
        n, d, k = int(rng.integers(5, 200)), int(rng.integers(1, 9)), int(rng.integers(1, 12))
        x = rng.normal(size=(n, d)) * 10.0 ** rng.integers(-4, 6, size=d)
        x[rng.random((n, d)) < 0.1] = np.nan
        codes = rng.integers(0, k, size=n)
        df = pd.DataFrame(x, columns=[f"c{j}" for j in range(d)])
        df["gid"] = codes
        want = df.groupby("gid").agg(["mean", "std"])
        present = np.array(sorted(set(codes.tolist())))
        mean, std = session_agg.group_mean_std(x, codes, k)
        for j in range(d):
            assert np.array_equal(mean[present, j], want[(f"c{j}", "mean")].to_numpy(), equal_nan=True)
            assert np.array_equal(std[present, j], want[(f"c{j}", "std")].to_numpy(), equal_nan=True)


class _FakeAggExtractor:
    def aggregate_sessions(self, x, codes, n_groups):
        from oracle import session_agg          # the checker standing in for the CUDA call (test only)
        return session_agg.group_mean_std(np.asarray(x, dtype=np.float64), np.asarray(codes), n_groups)


def test_aggregate_clip_features_contract(monkeypatch, capsys):
    """Host logic of the drop-in (merge, id order, column naming, empty input) without a GPU."""
    import pandas as pd
    from robust_speech_analysis_framework_b200 import mshds_extractor as mx, utils as bu
    monkeypatch.setattr(mx, "get_extractor", lambda device=0: _FakeAggExtractor())
    g, clip, meta = _agg_golden_frames()
    out = bu.aggregate_clip_features(clip, meta)
    assert list(out.columns) == ["unique_participant_id"] + list(g["out_columns"])
    assert list(out["unique_participant_id"]) == list(g["out_ids"])
    assert np.array_equal(out.iloc[:, 1:].to_numpy(dtype=np.float64), g["out_values"], equal_nan=True)
    empty = bu.aggregate_clip_features(pd.DataFrame(), meta)
    assert empty.empty and "Warning: Input clip_features_df is empty" in capsys.readouterr().out      # src/utils.py:31-33
    sys.path.insert(0, ROOT)
    from src.utils import aggregate_clip_features as shim
    assert shim is bu.aggregate_clip_features


# ------------------------------------------------------------------------------------------------ OpenSMILE LLD slice (8f-1)
def test_lld_oracle_known_answers():
    """Closed-form checks of the numpy restatement (oracle/lld_oracle.py): framing, ZCR, energy, mel bank, DCT."""
    from oracle import lld_oracle as lo
    fs = 16000.0
    # frame count: only complete 400-sample frames every 160 samples
    raw = dict(smooth_win=0, delta_win=0)
    assert lo.frame_lld(np.zeros(399), fs, **raw).shape == (0, 14) and lo.frame_lld(np.zeros(399), fs).shape == (0, 28)
    assert lo.frame_lld(np.zeros(400), fs, **raw).shape == (1, 14)
    assert lo.frame_lld(np.zeros(400 + 159), fs, **raw).shape == (1, 14) and lo.frame_lld(np.zeros(560), fs).shape == (2, 28)
    # a 1 kHz sine: 2 sign changes per period -> zcr ~ 2 * 1000 / 16000; energy of the pre-emphasised, windowed frame
    t = np.arange(16000) / fs
    x = 0.5 * np.sin(2 * np.pi * 1000.0 * t + 0.3)
    f = lo.frame_lld(x, fs, **raw)
    assert abs(f[:, 13].mean() - 2 * 1000.0 / fs) < 2.0 / 399
    h2 = 1.0 + 0.97 ** 2 - 2 * 0.97 * np.cos(2 * np.pi * 1000.0 / fs)          # |1 - k e^{-jw}|^2
    w = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(400) / 399)
    want = 0.5 * np.sqrt(h2 / 2.0) * np.sqrt((w * w).mean())
    assert abs(f[5:-5, 12].mean() / want - 1.0) < 2e-2
    # digital silence: log floor everywhere -> all cepstral coefficients of order >= 1 vanish (DCT of a constant)
    z = lo.frame_lld(np.zeros(800), fs, **raw)
    assert np.allclose(z[:, :12], 0.0, atol=1e-9) and np.all(z[:, 12:] == 0.0)
    # more mel bands / larger transform change the shapes consistently (BASELINE.json configs[4] sweep)
    for n_fft, n_mel in ((512, 40), (1024, 80), (2048, 40)):
        g = lo.frame_lld(x[:4000], fs, n_fft=n_fft, n_mel=n_mel)
        assert g.shape == (23, 28) and np.all(np.isfinite(g))
    fun, rows = lo.extract((x * 32767).astype(np.int16), np.array([0, 8000, 8000, 16000]), fs)
    assert fun.shape == (3, 56) and np.all(np.isnan(fun[1])) and len(rows[1]) == 0
    # smoothing and deltas: a linear ramp keeps its values under a symmetric moving average (away from the ends) and has a
    # constant regression delta equal to its slope
    ramp = np.arange(20.0)[:, None] * np.array([[1.0, -2.0]])
    sd = lo.smooth_delta(ramp, 3, 2)
    assert np.allclose(sd[3:-3, :2], ramp[3:-3]) and np.allclose(sd[3:-3, 2:], [[1.0, -2.0]])
    assert np.allclose(sd[0, :2], (2 * ramp[0] + ramp[1]) / 3)                 # the end frame is repeated


def test_lld_dataframe_contract(tmp_path, monkeypatch):
    import pandas as pd
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200 import lld_extractor as lx, mshds_extractor as mx

    class Fake:
        def lld_extract(self, pcm, offs, fs, want_frames=False, **params):
            fun, rows = lo.extract(pcm, offs, float(fs), **params)     # the checker standing in for the CUDA call (test only)
            return fun, None, None
    monkeypatch.setattr(mx, "get_extractor", lambda device=0: Fake())
    rng = np.random.default_rng(0)
    a = (rng.normal(scale=2000, size=8000)).astype(np.int16)
    pa, pb = str(tmp_path / "a.wav"), str(tmp_path / "b44.wav")
    _write_wav(pa, a)
    _write_wav(pb, a, fs=44100)
    frame = pd.DataFrame({"filepath": [pa, str(tmp_path / "nope.wav"), pb]})
    df = lx.extract_lld_functionals(frame, verbose=False, descriptor_set=0, functional_set=0)      # the first slice
    assert list(df.columns) == ["filename"] + lx.functional_names() and df.shape == (3, 57)
    assert df.iloc[1, 1:].isna().all() and not df.iloc[0, 1:].isna().any() and not df.iloc[2, 1:].isna().any()
    assert df.columns[1] == "mfcc_sma[1]_amean" and df.columns[-1] == "pcm_zcr_sma_de_stddev"
    assert "pcm_RMSenergy_sma_amean" in df.columns and "mfcc_sma_de[12]_stddev" in df.columns
    full = lx.extract_lld_functionals(frame, verbose=False)                                         # default: the widest set
    assert full.shape == (3, 721) and list(full.columns[1:]) == lx.functional_names(descriptor_set=1, functional_set=1)
    assert full.iloc[1, 1:].isna().all() and not full.iloc[0, 1:].isna().any()
    assert np.allclose(full["mfcc_sma[1]_amean"].values[[0, 2]], df["mfcc_sma[1]_amean"].values[[0, 2]])


def test_reference_probe_switches_to_the_real_extractor_when_it_is_importable(orc):
    """when praat-parselmouth is importable the unmodified reference function becomes the checker and
    the CPU baseline (bench.py, make_golden.py --from-reference); otherwise the probe says why and the port is used."""
    from oracle import reference_probe as probe
    if not probe.available():
        assert "parselmouth" in probe.why_not()
        return
    # un-circular parity: the CPU port against Praat itself on the golden clips
    g = np.load(GOLDEN)
    ref, cols = probe.extract(g["pcm"], g["offsets"], 16000)
    assert cols == list(g["feature_names"])
    port, _ = orc.extract(g["pcm"], g["offsets"], 16000.0)
    assert np.array_equal(np.isnan(ref), np.isnan(port))
    np.testing.assert_allclose(port, ref, rtol=1e-4, atol=1e-6, equal_nan=True)


def test_appendix_c_switches_default_to_the_golden_reading_and_move_only_their_columns(orc):
    """Every unverifiable Praat detail sits behind an oracle switch (praat_core.h OrcOptions).  Defaults reproduce the goldens
    (checked by test_oracle_reproduces_golden); a flipped switch may only move the columns of its own feature group."""
    from robust_speech_analysis_framework_b200.synth import synth_clip
    clip = synth_clip(100, 3.0).numpy()
    off = np.array([0, len(clip)], np.int64)
    orc.reset_options()
    base, _ = orc.extract(clip, off, 16000.0)
    allowed = {"silence_boundary": range(0, 5), "cut_interval": range(0, 5), "theil_tilt_complete": [11], "theil_cpps_complete": [12],
               "cpps_fit_range": [12], "cpps_time_frames": [12], "cpps_smooth_align": [12], "vuv_overlap": [12], "ltas_fill": [10, 11],
               "candidate_bound": list(range(0, 7)) + [9, 10, 11, 12] + list(range(13, 25))}
    try:
        for name, alts in orc.OPTIONS.items():
            for v in alts:
                orc.reset_options()
                orc.set_option(name, v)
                got, _ = orc.extract(clip, off, 16000.0)
                changed = [k for k in range(25) if not (got[0, k] == base[0, k] or (np.isnan(got[0, k]) and np.isnan(base[0, k])))]
                assert set(changed) <= set(allowed[name]), (name, v, changed)
        with pytest.raises(KeyError):
            orc.set_option("no_such_switch", 1)
    finally:
        orc.reset_options()
    again, _ = orc.extract(clip, off, 16000.0)
    assert np.array_equal(again, base, equal_nan=True)


def test_bench_row_comparison_flags_decision_flips():
    import bench
    rng = np.random.default_rng(0)
    want = rng.uniform(1, 5, size=(4, 25))
    got = want.copy()
    r = bench.compare_rows(got, want)
    assert r["parity_ok"] and r["decision_flips"] == 0 and r["max_rel_diff"] == 0.0
    got[2, 0] *= 1.01            # one syllable more in a speech-rate column
    got[1, 20] *= 1 + 1e-8       # within the continuous tolerance
    r = bench.compare_rows(got, want)
    assert not r["parity_ok"] and r["decision_flips"] == 1 and r["columns_out_of_tolerance"] == [0]
    got = want.copy(); got[0, 5] = np.nan
    assert bench.compare_rows(got, want)["nan_mismatches"] == 1


def test_register_fft_core_against_direct_dft(tmp_path):
    """csrc/fftreg.cuh compiles for the host: the register FFTs, the pass-A / exchange / pass-B split and the shuffle untangle
    (emulated lane by lane) reproduce a direct DFT / direct autocorrelation to rounding error."""
    import subprocess
    exe = str(tmp_path / "fftreg_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "fftreg_host_test.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_lld_oracle_functionals_and_spectral_descriptors_known_answers():
    """oracle/lld_oracle.py second slice against independent formulas: the twelve functionals vs numpy / scipy, spectral
    descriptors of a pure tone (centroid and roll-off at the tone, tiny flatness, flux 0 for a stationary signal), and the
    OpenSMILE-style column names of the host mirror."""
    import scipy.stats as st
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200.lld_extractor import functional_names
    rng = np.random.default_rng(11)
    y = rng.normal(size=(200, 3)) + np.linspace(0, 2, 200)[:, None] * np.array([1.0, -0.5, 0.0])[None, :]
    F = lo.functionals12(y)
    t = np.arange(200.0)
    for c in range(3):
        m, b = np.polyfit(t, y[:, c], 1)
        np.testing.assert_allclose(F[6:8, c], [m, b], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(F[8, c], np.mean((y[:, c] - (m * t + b)) ** 2), rtol=1e-9)
        np.testing.assert_allclose(F[9:12, c], [y[:, c].std(), st.skew(y[:, c]), st.kurtosis(y[:, c], fisher=False)], rtol=1e-9)
        assert F[3, c] == np.argmax(y[:, c]) and F[4, c] == np.argmin(y[:, c]) and F[2, c] == F[0, c] - F[1, c]
        np.testing.assert_allclose(F[5, c], y[:, c].mean(), rtol=1e-12)
    fs = 16000.0
    x = 0.5 * np.sin(2 * np.pi * 2000.0 * np.arange(int(0.5 * fs)) / fs)
    rows = lo.frame_lld(x, fs, descriptor_set=1, smooth_win=0, delta_win=0)
    q = rows[5:-5, 14:]
    assert np.all(np.abs(q[:, 9] - 2000.0) < 40.0) and np.all(np.abs(q[:, 5] - 2000.0) <= fs / 512)      # centroid, 50 % roll-off
    assert np.all(q[:, 15] < 1e-3) and np.all(q[:, 8] < 1e-2 * np.sqrt(q[:, 0] * 1e-6))                     # flatness, flux
    assert np.all(q[:, 3] > 100 * q[:, 2])                                                                    # 1-4 kHz band holds the tone
    np.testing.assert_allclose(q[:, 1], q[:, 0] ** 0.3, rtol=1e-12)
    names = functional_names(descriptor_set=1, functional_set=1)
    assert len(names) == 720 and names[0] == "mfcc_sma[1]_max" and "pcm_fftMag_spectralRollOff90.0_sma_de_kurtosis" in names
    assert len(functional_names()) == 56 and functional_names()[0] == "mfcc_sma[1]_amean"


def test_opensmile_drop_in_entry_point_and_config_parser(tmp_path, monkeypatch, capsys):
    """extract_opensmile_features keeps the reference's signature and result shape (src/opensmile_extractor.py:9-103: feature
    columns then 'filename', failed files left out, empty frame + warning when nothing worked) and takes its settings from the
    OpenSMILE configuration file."""
    import pandas as pd
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200 import lld_extractor as lx, mshds_extractor as mx
    from src.opensmile_extractor import extract_opensmile_features

    conf = tmp_path / "Androids_fixed.conf"
    conf.write_text("""
[componentInstances:cComponentManager]
instance[waveIn].type=cWaveSource
[fr1:cFramer]
reader.dmLevel=wave
frameSize=0.0250   ; seconds
frameStep = 0.010
[pe2:cVectorPreemphasis]
k=0.95
[mspec:cMelspec]
htkcompatible = 1
lofreq = 50   // Hz
hifreq = 7000
[mfcc:cMfcc]
firstMfcc = 1
lastMfcc =  10
[shs:cPitchShs]
maxPitch = 620
[pitchJitter:cPitchJitter]
jitterLocal = 1
[delta1:cDeltaRegression]
deltawin=2
[functL1:cFunctionals]
frameSize=0.025
frameStep=0
""")
    params, missing = lx.parse_smile_config(str(conf))
    assert params == {"frame_size": 0.025, "frame_step": 0.01, "preemph": 0.95, "mel_lo": 50.0, "mel_hi": 7000.0, "n_mfcc": 10,
                      "delta_win": 2}
    assert missing == ["cPitchShs", "cPitchJitter"]
    seen = {}

    class Fake:
        def lld_extract(self, pcm, offs, fs, want_frames=False, **p):
            seen.update(p)
            fun, _ = lo.extract(pcm, offs, float(fs), **p)           # the checker standing in for the CUDA call (test only)
            return fun, None, None
    monkeypatch.setattr(mx, "get_extractor", lambda device=0: Fake())
    rng = np.random.default_rng(3)
    pa = str(tmp_path / "a.wav")
    _write_wav(pa, (rng.normal(scale=2000, size=8000)).astype(np.int16))
    frame = pd.DataFrame({"filepath": [pa, str(tmp_path / "missing.wav")]})
    df = extract_opensmile_features(frame, "C:/tools/opensmile/bin/SMILExtract.exe", str(conf), verbose=True)
    out = capsys.readouterr().out
    assert "cPitchShs" in out and "ERROR processing missing.wav" in out
    assert seen["n_mfcc"] == 10 and seen["preemph"] == 0.95 and seen["descriptor_set"] == 1 and seen["functional_set"] == 1
    assert df.shape == (1, (10 + 2 + 16) * 2 * 12 + 1) and df.columns[-1] == "filename" and df["filename"].iloc[0] == "a.wav"
    assert df.columns[0] == "mfcc_sma[1]_max" and not df.iloc[0, :-1].isna().any()
    empty = extract_opensmile_features(frame.iloc[1:], None, None, verbose=False)
    assert empty.empty and "No features were successfully extracted" in capsys.readouterr().out


def test_sliding_sum_cross_correlation_is_exact_in_float64():
    """The arithmetic claim behind k_cc_frames_w / k_cc_frames_s (k_ccs.cu): for int16 samples s = v / 32768 every product
    s[j] s[j + lag] is an integer multiple of 2^-30 and window sums stay far below 2^53, so adding the products that enter a
    window and subtracting those that leave it is EXACT in float64 -- no drift over a run of frames, whatever the order.  Checked
    against Python integers: 400 frames of the harmonicity geometry (W = 1198, step 80, 268 lags), full-scale noise + DC."""
    rng = np.random.default_rng(12)
    W, H, L, nfr = 1198, 80, 268, 400
    v = (rng.integers(-32768, 32768, size=W + H * nfr + L + 8)).astype(np.int64)
    v[5000:9000] = 32767                                           # a full-scale DC stretch: the largest possible sums
    s = v.astype(np.float64) / 32768.0
    lags = np.array([1, 2, 17, 160, 267, 268])
    C = np.array([np.dot(s[:W], s[l:l + W]) for l in lags])        # frame 0 (any summation order: exact as well)
    Ci = [int(np.dot(v[:W], v[l:l + W])) for l in lags]
    for k in range(1, nfr + 1):
        lo, hi = (k - 1) * H, (k - 1) * H + W                      # [lo, lo + H) leaves, [hi, hi + H) enters
        for n, l in enumerate(lags):
            add = 0.0
            for j in range(hi, hi + H):
                add = add + s[j] * s[j + l]
            sub = 0.0
            for j in range(lo, lo + H):
                sub = sub + s[j] * s[j + l]
            C[n] = C[n] + add - sub
            Ci[n] += int(np.dot(v[hi:hi + H], v[hi + l:hi + H + l])) - int(np.dot(v[lo:lo + H], v[lo + l:lo + H + l]))
    direct = [int(np.dot(v[nfr * H:nfr * H + W], v[nfr * H + l:nfr * H + l + W])) for l in lags]
    assert Ci == direct
    assert [float(c) for c in C] == [ci / float(1 << 30) for ci in Ci]            # bit-exact, not approximately equal
    assert max(abs(ci) for ci in Ci) < (1 << 53)


def test_lld_oracle_reproduces_its_golden_vectors():
    """tests/golden/lld_golden_v1.npz freezes oracle/lld_oracle.py (make_lld_golden.py): the restatement cannot drift silently.
    Tolerance 1e-9 relative, absolute 1e-9 of each column's largest value (the numpy / BLAS build may reorder sums)."""
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200.lld_extractor import functional_names
    g = np.load(os.path.join(ROOT, "tests", "golden", "lld_golden_v1.npz"))
    pcm, off = g["pcm"], g["offsets"]
    fun56, _ = lo.extract(pcm, off, 16000.0)
    np.testing.assert_allclose(fun56, g["functionals_56"], rtol=1e-9, atol=1e-9, equal_nan=True)
    fun, rows = lo.extract(pcm, off, 16000.0, descriptor_set=1, functional_set=1)
    frames = np.concatenate([r for r in rows if len(r)])
    assert [len(r) for r in rows] == list(g["frame_counts"])
    sc = np.abs(g["frames_720"]).max(axis=0) + 1e-300
    assert np.all(np.abs(frames - g["frames_720"]) <= 1e-9 * np.abs(g["frames_720"]) + 1e-9 * sc[None, :])
    assert np.array_equal(np.isnan(fun), np.isnan(g["functionals_720"]))
    assert list(g["names_720"]) == functional_names(descriptor_set=1, functional_set=1) and list(g["names_56"]) == functional_names()
    ok = ~np.isnan(fun)
    W = frames.shape[1]
    pos = np.zeros(12 * W, bool); pos[3 * W:5 * W] = True                 # maxPos / minPos: indices, exact
    assert np.array_equal(fun[:, pos], g["functionals_720"][:, pos], equal_nan=True)
    np.testing.assert_allclose(fun[ok], g["functionals_720"][ok], rtol=1e-6, atol=1e-6 * float(np.abs(g["functionals_720"][ok]).max()))
