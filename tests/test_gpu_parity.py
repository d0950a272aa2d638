"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI (ctypes binding of
include/mshds_b200.h), against the CPU oracle and the frozen golden vectors.

Tolerances (BASELINE.md / north_star): spectral, LTAS, CPP, intensity, HNR columns <= 1e-4 relative; mean_F0 <= 0.5 Hz;
formant statistics <= 1e-3 relative; count-derived speech-rate columns exact.  The kernels do everything decision-critical
in float64, so the tests hold them to much tighter bounds (written next to each assert).
"""
import os
import wave

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "mshds_golden_v1.npz")

SPEECHRATE = slice(0, 5)
REL_TOL = {  # per column group: (rtol, atol)
    "speechrate": (1e-12, 1e-12),      # decisions agree exactly; only the final divisions could differ in the last bit
    "pitch": (1e-7, 1e-7),             # north_star: 0.5 Hz
    "continuous": (1e-6, 1e-9),        # north_star: 1e-4 relative
    "formant": (1e-5, 1e-6),           # north_star: 1e-3 relative (different root finders: Aberth vs Hessenberg QR)
}
GROUP_OF = ["speechrate"] * 5 + ["pitch"] * 2 + ["continuous"] * 6 + ["formant"] * 8 + ["continuous"] * 4


@pytest.fixture(scope="module")
def ex():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from robust_speech_analysis_framework_b200 import _lib
    e = _lib.Extractor(0)
    yield e
    e.close()


def assert_features_close(got, want, what=""):
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want)), f"{what}: NaN pattern differs\n{got}\n{want}"
    for k in range(25):
        rtol, atol = REL_TOL[GROUP_OF[k]]
        np.testing.assert_allclose(got[:, k], want[:, k], rtol=rtol, atol=atol, equal_nan=True, err_msg=f"{what} column {k}")


def test_golden_vectors(ex):
    g = np.load(GOLDEN)
    got, status = ex.extract_host(g["pcm"], g["offsets"])
    assert_features_close(got, g["features"], "golden")
    assert np.array_equal(status, g["status"])


def _batch(durs, start=0):
    from robust_speech_analysis_framework_b200.synth import synth_clip
    clips = [synth_clip(start + i, d).numpy() for i, d in enumerate(durs)]
    return np.concatenate(clips), np.cumsum([0] + [len(c) for c in clips]).astype(np.int64), clips


def test_parity_with_oracle_on_ragged_batch(ex, orc):
    pcm, off, _ = _batch([6.0, 2.00006, 9.3, 4.1, 12.0, 3.33339], start=20)
    got, st = ex.extract_host(pcm, off)
    want, wst = orc.extract(pcm, off, 16000.0, nthreads=os.cpu_count() or 1)
    assert_features_close(got, want, "ragged batch")
    assert np.array_equal(st, wst)
    # the count-derived columns must be bit-identical unless a decision flipped (none may)
    assert np.array_equal(got[:, SPEECHRATE], want[:, SPEECHRATE])


def test_parity_on_stress_clips_with_short_bursts_and_gaps(ex, orc):
    """"stress" clips (synth.py): sounding / silent intervals below the 0.1 s / 0.3 s minima that the silence detector must cut
    and merge, voiced runs whose +-50 ms extensions overlap, 30 % unvoiced syllables -- the decisions the plain clips never
    reach (tests/golden/appc_sensitivity.md shows they move Pause_Rate and CPP when read differently)."""
    from robust_speech_analysis_framework_b200.synth import synth_clip
    clips = [synth_clip(700 + i, d, style="stress").numpy() for i, d in enumerate([8.0, 5.00006, 11.0, 6.5])]
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    got, st = ex.extract_host(pcm, off)
    want, wst = orc.extract(pcm, off, 16000.0, nthreads=os.cpu_count() or 1)
    assert_features_close(got, want, "stress clips")
    assert np.array_equal(st, wst)
    assert np.array_equal(got[:, SPEECHRATE], want[:, SPEECHRATE])


def test_stage_level_parity(ex, orc):
    pcm, off, clips = _batch([5.0, 4.00006], start=40)
    ex.extract_host(pcm, off)
    for c, clip in enumerate(clips):
        x = orc.pcm_to_float(clip)
        fl, ce, _ = orc.pitch_values(x)
        cls = int(ex.debug_fetch("class", c, np.int32)[0])
        assert (60.0, 100.0, 75.0)[cls] == fl
        ref = orc.pitch(x, 16000.0, 0, 0.005, fl, 3.0, 15, 0.03, 0.45, 0.01, 0.35, 0.14, ce)
        got_f = ex.debug_fetch("pitch_main_f", c)
        assert len(got_f) == len(ref["freq"])
        assert np.array_equal(got_f > 0, ref["freq"] > 0)                       # voicing decisions identical
        np.testing.assert_allclose(got_f, ref["freq"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(ex.debug_fetch("pitch_main_s", c), ref["strength"], rtol=0, atol=1e-9)
        ic, _ = orc.intensity(x, 16000.0, fl, 0.005)
        np.testing.assert_allclose(ex.debug_fetch("intensity_main", c), ic, rtol=0, atol=1e-10)
        isr, _ = orc.intensity(x, 16000.0, 50.0, 0.016)
        np.testing.assert_allclose(ex.debug_fetch("intensity_sr", c), isr, rtol=0, atol=1e-10)
        for name, meth, ppw, vt, dt in (("pulses_ltas", 0, 3.0, 0.45, 0.0), ("pulses_cpp", 0, 3.0, 0.3, 0.005), ("pulses_fmt", 2, 1.0, 0.45, 0.005)):
            want_p = orc.pulses(x, 16000.0, meth, dt, fl, ppw, vt, ce)
            got_p = ex.debug_fetch(name, c)
            assert len(got_p) == len(want_p), name
            np.testing.assert_allclose(got_p, want_p, rtol=0, atol=1e-9, err_msg=name)
        rs, _ = orc.resample(x, 16000.0, 10000.0, 500)
        np.testing.assert_allclose(ex.debug_fetch("resampled10k", c), rs, rtol=0, atol=1e-10)
        fo = orc.formants(x, 16000.0)
        assert np.array_equal(ex.debug_fetch("formant_n", c, np.int32), fo["n"])
        np.testing.assert_allclose(ex.debug_fetch("formant_f", c).reshape(-1, 5), fo["f"], rtol=1e-6, atol=1e-4, equal_nan=True)
        bands, _ = orc.ltas(x, 16000.0, fl, ce)
        np.testing.assert_allclose(ex.debug_fetch("ltas_bands", c), bands, rtol=0, atol=1e-9)


def test_edge_cases_follow_reference_error_convention(ex, orc):
    rng = np.random.default_rng(3)
    clips = [
        np.zeros(0, np.int16),                                   # empty recording -> whole row NaN (:450-457)
        np.zeros(32000, np.int16),                               # digital silence
        np.full(20000, 1234, np.int16),                          # DC only
        (rng.normal(scale=300, size=100)).astype(np.int16),      # shorter than every window
        (rng.normal(scale=500, size=1500)).astype(np.int16),     # ~0.094 s: some analyses fit, others throw
        (rng.normal(scale=500, size=2300)).astype(np.int16),     # ~0.144 s
        (rng.normal(scale=900, size=40000)).astype(np.int16),    # unvoiced noise
    ]
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    got, st = ex.extract_host(pcm, off)
    assert np.all(np.isnan(got[0])) and st[0] == (1 << 31)
    want, wst = orc.extract(pcm, off, 16000.0, nthreads=4)
    assert_features_close(got, want, "edge cases")
    assert np.array_equal(st, wst)


def test_batch_composition_does_not_change_a_clip(ex):
    pcm, off, clips = _batch([3.0, 5.5, 2.2, 4.4, 3.3], start=60)
    full, _ = ex.extract_host(pcm, off)
    # permuted order
    perm = [3, 0, 4, 2, 1]
    p2 = np.concatenate([clips[i] for i in perm])
    o2 = np.cumsum([0] + [len(clips[i]) for i in perm]).astype(np.int64)
    permuted, _ = ex.extract_host(p2, o2)
    assert np.array_equal(permuted, full[perm], equal_nan=True)              # bit-identical
    # one clip alone, and duplicated
    alone, _ = ex.extract_host(clips[2], np.array([0, len(clips[2])], np.int64))
    assert np.array_equal(alone[0], full[2], equal_nan=True)
    dup, _ = ex.extract_host(np.concatenate([clips[2], clips[2]]), np.array([0, len(clips[2]), 2 * len(clips[2])], np.int64))
    assert np.array_equal(dup[0], dup[1], equal_nan=True) and np.array_equal(dup[0], full[2], equal_nan=True)
    # chunked execution (scratch bounded to ~2 clips at a time) gives the same bits
    ex.set_chunk_samples(100000)
    try:
        chunked, _ = ex.extract_host(pcm, off)
    finally:
        ex.set_chunk_samples(1 << 27)
    assert np.array_equal(chunked, full, equal_nan=True)


def test_device_resident_entry_point(ex):
    import torch
    pcm, off, _ = _batch([2.5, 3.5], start=80)
    host, hst = ex.extract_host(pcm, off)
    d_pcm = torch.from_numpy(pcm).cuda()
    d_out = torch.empty((2, 25), dtype=torch.float64, device="cuda")
    d_st = torch.empty(2, dtype=torch.int32, device="cuda")
    ex.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        ex.extract_device(d_pcm.data_ptr(), off, d_out.data_ptr(), d_st.data_ptr())
        torch.cuda.synchronize()
    finally:
        ex.reset_stream()
    assert np.array_equal(d_out.cpu().numpy(), host, equal_nan=True)
    assert np.array_equal(d_st.cpu().numpy().astype(np.uint32), hst)


def test_caller_stream_orders_the_extraction_after_async_producers(ex):
    """the int16 batch is PRODUCED asynchronously on a non-default torch stream (behind a long-running kernel);
    the extraction is handed that stream and must see the finished data -- and likewise on the legacy default stream,
    which is what a NULL cudaStream_t names."""
    import torch
    pcm, off, _ = _batch([2.5, 3.0], start=84)
    host, hst = ex.extract_host(pcm, off)
    src = torch.from_numpy(pcm).cuda()
    big = torch.empty(1 << 28, dtype=torch.float32, device="cuda")          # 1 GiB: keeps the stream busy for a while
    for stream in (torch.cuda.Stream(), None):
        d_pcm = torch.zeros_like(src)
        d_out = torch.full((2, 25), -1.0, dtype=torch.float64, device="cuda")
        d_st = torch.empty(2, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.default_stream())
        with ctx:
            for _ in range(8):
                big.mul_(1.0001)                                            # queued work in front of the producer
            d_pcm.copy_(src, non_blocking=True)                             # the producer, asynchronous on this stream
            ex.set_stream(torch.cuda.current_stream().cuda_stream)          # 0 for the default stream
            try:
                ex.extract_device(d_pcm.data_ptr(), off, d_out.data_ptr(), d_st.data_ptr())
            finally:
                ex.reset_stream()
        torch.cuda.synchronize()
        assert np.array_equal(d_out.cpu().numpy(), host, equal_nan=True), "extraction read the batch before it was written"
        assert np.array_equal(d_st.cpu().numpy().astype(np.uint32), hst)


def test_warp_resident_fft_frames_equal_the_shared_memory_ones(ex):
    """Round 2 frame kernel (k_acw.cu: warp-per-frame register FFT, TMA-staged sample spans) against the round-1
    CTA-per-frame shared-memory FFT kernel ("legacy_fft"): same correlation up to the rounding of another exact FFT
    order, so identical voicing decisions and F0 contours to 1e-9 on both speaker classes and on odd-length clips."""
    pcm, off, clips = _batch([4.0, 3.00006, 2.5, 5.1], start=60)
    new, nst = ex.extract_host(pcm, off)
    cn = {k: [ex.debug_fetch(k, c) for c in range(len(clips))] for k in ("pitch_wide_f", "pitch_main_f", "pitch_main_s", "pitch_ltas_f", "pitch_cpp_f")}
    ex.set_option("legacy_fft", 1)
    try:
        old, ost = ex.extract_host(pcm, off)
        co = {k: [ex.debug_fetch(k, c) for c in range(len(clips))] for k in cn}
    finally:
        ex.set_option("legacy_fft", 0)
    assert np.array_equal(nst, ost)
    for k in cn:
        for a, b in zip(cn[k], co[k]):
            assert len(a) == len(b)
            assert np.array_equal(a == 0, b == 0), f"{k}: voicing decisions differ"
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9, err_msg=k)   # Brent stops at sqrt(eps) |x|: 1.5e-8 relative
    assert_features_close(new, old, "warp FFT vs legacy FFT")
    assert np.array_equal(new[:, SPEECHRATE], old[:, SPEECHRATE])


def test_shared_block_cross_correlation_equals_frame_by_frame(ex, orc):
    """k_ccs.cu (block products shared between overlapping frames, exact for int16 input) against the frame-by-frame
    cross-correlation kernel ("legacy_cc") and the oracle: harmonicity contour, to_pitch_cc contour, formant pulses.  Clips of
    both speaker classes, odd lengths (frame centres on sample boundaries), a short clip (runs shorter than the ring), a clip
    with a large DC offset (the mean correction cancels most of the raw products) and digital silence inside a clip."""
    from robust_speech_analysis_framework_b200.synth import synth_clip
    clips = [synth_clip(300 + i, d).numpy() for i, d in enumerate([4.0, 3.00006, 0.35, 5.1, 2.2])]
    dc = synth_clip(310, 3.0).numpy().astype(np.int32) // 4 + 9000
    clips.append(dc.astype(np.int16))
    gap = synth_clip(311, 3.0).numpy().copy()
    gap[16000:24000] = 0
    clips.append(gap)
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    keys = ("hnr_r", "pitch_cc_f", "pulses_fmt")
    new, nst = ex.extract_host(pcm, off)
    cn = {k: [ex.debug_fetch(k, c) for c in range(len(clips))] for k in keys}
    ex.set_option("legacy_cc", 1)
    try:
        old, ost = ex.extract_host(pcm, off)
        co = {k: [ex.debug_fetch(k, c) for c in range(len(clips))] for k in keys}
    finally:
        ex.set_option("legacy_cc", 0)
    assert np.array_equal(nst, ost)
    # the CTA block-ring kernel (k_cc_frames_s, "legacy_cc" = 2) and the warp-per-run sliding sums (k_cc_frames_w, default) are
    # both exact on int16 input: bit-identical rows
    ex.set_option("legacy_cc", 2)
    try:
        ring, rst = ex.extract_host(pcm, off)
        cr = {k: [ex.debug_fetch(k, c) for c in range(len(clips))] for k in keys}
    finally:
        ex.set_option("legacy_cc", 0)
    assert np.array_equal(ring, new, equal_nan=True) and np.array_equal(rst, nst)
    for k in keys:
        for ci, (a, b) in enumerate(zip(cn[k], cr[k])):
            assert np.array_equal(a, b, equal_nan=True), f"{k} clip {ci}: sliding sums differ from the block ring"
    for k in keys:
        for ci, (a, b) in enumerate(zip(cn[k], co[k])):
            assert len(a) == len(b), (k, ci)
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a == 0, b == 0), f"{k} clip {ci}: decisions differ"
            np.testing.assert_allclose(a, b, rtol=1e-7, atol=1e-9, equal_nan=True, err_msg=f"{k} clip {ci}")
    assert_features_close(new, old, "shared-block CC vs frame-by-frame CC")
    assert np.array_equal(new[:, SPEECHRATE], old[:, SPEECHRATE])
    want, wst = orc.extract(pcm, off, 16000.0, nthreads=os.cpu_count() or 1)
    assert_features_close(new, want, "shared-block CC vs oracle")
    assert np.array_equal(nst, wst)
    # exact arithmetic on int16 products: the grouping of frames into runs cannot show in the result
    ex.set_chunk_samples(70000)
    try:
        chunked, _ = ex.extract_host(pcm, off)
    finally:
        ex.set_chunk_samples(1 << 27)
    assert np.array_equal(chunked, new, equal_nan=True)


def _sharded_worker(rank, world, port, pcm, off, q):
    import torch
    import torch.distributed as dist
    from robust_speech_analysis_framework_b200 import _lib
    from robust_speech_analysis_framework_b200.sharding import extract_sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    e = _lib.Extractor(rank)
    out, st = extract_sharded(pcm, off, e.extract_host, rank, world)
    if rank == 0:
        q.put((out, st))
    dist.barrier()
    dist.destroy_process_group()
    e.close()


def test_two_gpu_sharded_extraction_equals_single_gpu(ex):
    """LPT-sharded extraction over 2 GPUs (one process each, NCCL gather of the feature matrix only) is
    bit-identical to the single-GPU result, row for row in input order."""
    import socket
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    pcm, off, _ = _batch([6.0, 2.0, 9.0, 3.5, 4.00006, 7.0, 2.5], start=140)
    want, wst = ex.extract_host(pcm, off)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, pcm, off, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, gst = q.get(timeout=600)
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    assert np.array_equal(got, want, equal_nan=True) and np.array_equal(gst, wst)


def test_bad_arguments_are_errors_not_crashes(ex):
    from robust_speech_analysis_framework_b200 import _lib
    pcm = np.zeros(1000, np.int16)
    with pytest.raises(_lib.MshdsError):
        ex.extract_host(pcm, np.array([0, 1000], np.int64), sample_rate=100)
    with pytest.raises(_lib.MshdsError):
        ex.extract_host(pcm, np.array([500, 100], np.int64))
    out, st = ex.extract_host(np.zeros(0, np.int16), np.array([0], np.int64))
    assert out.shape == (0, 25)


def _write_wav(path, x, fs=16000):
    with wave.open(path, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(fs)
        w.writeframes(np.asarray(x, dtype="<i2").tobytes())


def test_drop_in_dataframe_api_feeds_the_svm_consumer(tmp_path, orc):
    """extract_mshds_features on WAV files, then the reference's consumer pipeline (cv_strategies.py:38-78:
    StandardScaler -> SelectKBest(f_classif) -> linear SVC, stratified 5-fold) runs unchanged on the frame."""
    import pandas as pd
    from sklearn.feature_selection import SelectKBest, f_classif
    from sklearn.model_selection import StratifiedKFold, cross_val_score
    from sklearn.pipeline import Pipeline
    from sklearn.preprocessing import StandardScaler
    from sklearn.svm import SVC
    from src.mshds_extractor import extract_mshds_features
    from robust_speech_analysis_framework_b200.synth import synth_clip

    paths, clips = [], []
    for i in range(12):
        x = synth_clip(200 + i, 2.5 + 0.25 * (i % 3)).numpy()
        p = str(tmp_path / f"{i:02d}_clip.wav")
        _write_wav(p, x)
        paths.append(p)
        clips.append(x)
    paths.insert(5, str(tmp_path / "missing.wav"))
    df = extract_mshds_features(pd.DataFrame({"filepath": paths}), verbose=False)
    assert df.shape == (13, 26) and df.iloc[5, 1:].isna().all()
    assert list(df["filename"])[:2] == ["00_clip.wav", "01_clip.wav"]
    want, _ = orc.extract(np.concatenate(clips), np.cumsum([0] + [len(c) for c in clips]).astype(np.int64), 16000.0, nthreads=4)
    got = df.drop(index=5).iloc[:, 1:].to_numpy(dtype=np.float64)
    assert_features_close(got, want, "dataframe api")
    X = df.drop(columns=["filename"])
    X = X.fillna(X.mean())                                     # notebooks/02_model_evaluation.ipynb:155
    y = np.array([i % 2 for i in range(13)])                  # even / odd synthetic speakers = low / high F0 class
    pipe = Pipeline([("scaler", StandardScaler()), ("select", SelectKBest(f_classif, k=10)), ("svm", SVC(kernel="linear"))])
    scores = cross_val_score(pipe, X, y, cv=StratifiedKFold(3, shuffle=True, random_state=42))
    assert scores.shape == (3,) and np.all(np.isfinite(scores))


@pytest.mark.parametrize("fs", [44100, 48000, 22050, 32000, 11025, 8000])
def test_front_end_resamples_to_16k_like_the_reference(ex, orc, fs):
    """mshds_extractor.py:418-419: recordings at another rate go through snd.resample(16000, 50) first.  44.1 kHz and
    22.05 kHz have long polyphase periods (160 / 320 phases), 48 and 32 kHz short ones (FIR kernel), 11.025 kHz up-samples
    (no low-pass), 8 kHz is the exact doubling Praat hands to Sound_upsample.  Lengths include odd ones so that the new time
    origin is not half a sample."""
    from robust_speech_analysis_framework_b200.synth import synth_clip
    durs = [3.0, 2.2 + 1.0 / fs, 4.1 + 3.0 / fs]
    clips = [synth_clip(300 + i, d, fs=fs).numpy() for i, d in enumerate(durs)]
    clips.append(np.zeros(0, np.int16))                      # empty recording inside the batch
    clips.append(clips[0][: fs // 50])                       # 20 ms: too short for most analyses
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    got, st = ex.extract_host(pcm, off, sample_rate=fs)
    want, wst = orc.extract(pcm, off, float(fs), nthreads=os.cpu_count() or 1)
    assert_features_close(got, want, f"front-end {fs} Hz")
    assert np.array_equal(st, wst)
    assert np.array_equal(got[:, SPEECHRATE], want[:, SPEECHRATE], equal_nan=True)
    # the resampled signal itself (stage read-back of the formant path's source is indirect; compare the 16 kHz sound)
    x16 = ex.debug_fetch("resampled16k", 1)
    ref, _x1 = orc.resample(orc.pcm_to_float(clips[1]), float(fs), 16000.0, 50)
    assert len(x16) == len(ref)
    # coefficient rows are shared by all samples of a phase, whose fractional positions differ by rounding (~1e-10 of a sample)
    np.testing.assert_allclose(x16, ref, rtol=0, atol=1e-11)


def test_mixed_rate_dataframe(tmp_path, orc):
    import pandas as pd
    from src.mshds_extractor import extract_mshds_features
    from robust_speech_analysis_framework_b200.synth import synth_clip
    rates = [16000, 44100, 16000, 48000, 44100]
    paths, want = [], []
    for i, fs in enumerate(rates):
        x = synth_clip(400 + i, 2.5, fs=fs).numpy()
        p = str(tmp_path / f"r{i}.wav")
        _write_wav(p, x, fs)
        paths.append(p)
        w, _ = orc.extract(x, np.array([0, len(x)], np.int64), float(fs))
        want.append(w[0])
    df = extract_mshds_features(pd.DataFrame({"filepath": paths}), verbose=False)
    assert_features_close(df.iloc[:, 1:].to_numpy(dtype=np.float64), np.stack(want), "mixed rates")


def test_size_independent_properties_on_a_larger_batch(ex):
    """64 clips x 10 s: duplicates give identical rows, every column is populated, values stay in physical ranges."""
    from robust_speech_analysis_framework_b200.synth import synth_batch
    pcm, off = synth_batch(64, 10.0, unique=16, start_index=300)
    out, st = ex.extract_host(pcm.numpy(), off.numpy())
    assert not np.isnan(out).any() and not st.any()
    for i in range(16, 64):
        assert np.array_equal(out[i], out[i % 16])
    assert np.all((out[:, 5] > 80) & (out[:, 5] < 260)) and np.all((out[:, 2] > 0.4) & (out[:, 2] <= 1.0))


def test_long_clip_exercises_the_multi_pass_fft(ex, orc):
    """A 70 s clip needs a 2^21-point low-pass FFT (two strided passes around the in-SM block kernel)."""
    from robust_speech_analysis_framework_b200.synth import synth_clip
    clip = synth_clip(400, 70.0).numpy()
    got, st = ex.extract_host(clip, np.array([0, len(clip)], np.int64))
    rs, _ = orc.resample(orc.pcm_to_float(clip), 16000.0, 10000.0, 500)
    np.testing.assert_allclose(ex.debug_fetch("resampled10k", 0), rs, rtol=0, atol=1e-10)
    want, wst = orc.extract(clip, np.array([0, len(clip)], np.int64), 16000.0)
    assert_features_close(got, want, "70 s clip")
    assert np.array_equal(st, wst)


def test_many_short_clips_like_config_4(ex, orc):
    """BASELINE.json configs[3] in miniature: 2 s clips (launch / packing stress); every row against the oracle."""
    from robust_speech_analysis_framework_b200.synth import synth_batch
    pcm, off = synth_batch(192, 2.0, start_index=500)
    pcm, off = pcm.numpy(), off.numpy()
    got, st = ex.extract_host(pcm, off)
    want, wst = orc.extract(pcm, off, 16000.0, nthreads=os.cpu_count() or 1)
    assert_features_close(got, want, "192 x 2 s")
    assert np.array_equal(st, wst)
    assert np.array_equal(got[:, SPEECHRATE], want[:, SPEECHRATE], equal_nan=True)


# ------------------------------------------------------------------------------------------------ session aggregation (8f-3)
def test_session_aggregation_is_bit_identical_to_the_reference(ex):
    """mshds_aggregate_sessions against vectors produced by the reference's own aggregate_clip_features
    (src/utils.py:7-58, tests/golden/make_agg_golden.py) and against the Python restatement on random groups."""
    import pandas as pd
    from oracle import session_agg
    from robust_speech_analysis_framework_b200 import utils as bu
    g = np.load(os.path.join(ROOT, "tests", "golden", "session_agg_golden_v1.npz"), allow_pickle=True)
    clip = pd.DataFrame(g["clip_values"], columns=list(g["clip_columns"]))
    clip.insert(0, "filename", list(g["clip_filenames"]))
    meta = pd.DataFrame({"filename": list(g["meta_filenames"]), "unique_participant_id": list(g["meta_ids"])})
    out = bu.aggregate_clip_features(clip, meta)
    assert list(out.columns) == ["unique_participant_id"] + list(g["out_columns"])
    assert list(out["unique_participant_id"]) == list(g["out_ids"])
    assert np.array_equal(out.iloc[:, 1:].to_numpy(dtype=np.float64), g["out_values"], equal_nan=True)     # bit for bit
    rng = np.random.default_rng(11)
    for n, d, k in ((1, 1, 1), (7, 3, 9), (866, 25, 109), (5000, 50, 400)):
        x = rng.normal(size=(n, d)) * 10.0 ** rng.integers(-4, 6, size=d)
        x[rng.random((n, d)) < 0.08] = np.nan
        codes = rng.integers(-1, k, size=n).astype(np.int32)            # -1 = row without a session
        mean, std = ex.aggregate_sessions(x, codes, k)
        wm, ws = session_agg.group_mean_std(x, codes, k)
        assert np.array_equal(mean, wm, equal_nan=True) and np.array_equal(std, ws, equal_nan=True)
    from robust_speech_analysis_framework_b200 import _lib
    with pytest.raises(_lib.MshdsError):
        ex.aggregate_sessions(np.zeros((3, 2)), np.array([0, 5, 1], np.int32), 2)


def test_session_aggregation_on_device_buffers(ex):
    """Extraction output stays in HBM and is aggregated there; only the group list crosses PCIe."""
    import torch
    from oracle import session_agg
    pcm, off, _ = _batch([2.0, 2.5, 3.0, 2.2, 2.7], start=70)
    dpcm = torch.from_numpy(pcm).cuda()
    feats = torch.empty((5, 25), dtype=torch.float64, device="cuda")
    status = torch.empty(5, dtype=torch.int32, device="cuda")
    ex.extract_device(dpcm.data_ptr(), off, feats.data_ptr(), status.data_ptr())
    codes = np.array([1, 0, 1, 1, 0], np.int32)
    mean = torch.empty((2, 25), dtype=torch.float64, device="cuda")
    std = torch.empty((2, 25), dtype=torch.float64, device="cuda")
    ex.aggregate_sessions_device(feats.data_ptr(), 5, 25, codes, 2, mean.data_ptr(), std.data_ptr())
    torch.cuda.synchronize()
    wm, ws = session_agg.group_mean_std(feats.cpu().numpy(), codes, 2)
    assert np.array_equal(mean.cpu().numpy(), wm, equal_nan=True) and np.array_equal(std.cpu().numpy(), ws, equal_nan=True)


def test_androids_scale_ragged_batch_against_stored_oracle_values(ex):
    """BASELINE.json configs[2]: 1-10 min recordings of ragged length (one 10 min clip = a 2^24-point resampling FFT and
    120k frames per pitch pass).  The CPU oracle needs minutes for these, so its values were computed once in the build
    container (tests/golden/make_long_golden.py); the clips are re-synthesised here from their seeds."""
    import hashlib
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_long_golden", os.path.join(ROOT, "tests", "golden", "make_long_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(ROOT, "tests", "golden", "long_clips_golden_v1.npz"))
    pcm, off = mod.make_batch()
    if hashlib.sha256(pcm.tobytes()).hexdigest() != str(g["sha256"]):
        pytest.skip("synthetic clips differ from the ones the stored oracle values were computed on")
    got, st = ex.extract_host(pcm, off)
    assert_features_close(got, g["features"], "1-10 min ragged batch")
    assert np.array_equal(st, g["status"])
    assert np.array_equal(got[:, SPEECHRATE], g["features"][:, SPEECHRATE])
    # the 10 min clip alone, in a chunk of its own, gives the same row bit for bit (batch composition invariance)
    alone, _ = ex.extract_host(pcm[off[0]:off[1]], np.array([0, off[1] - off[0]], np.int64))
    assert np.array_equal(alone[0], got[0], equal_nan=True)


# ------------------------------------------------------------------------------------------------ OpenSMILE LLD slice (8f-1)
@pytest.mark.parametrize("fs,params", [
    (16000, {}),
    (16000, {"n_fft": 512, "n_mel": 40}), (16000, {"n_fft": 1024, "n_mel": 80}), (16000, {"n_fft": 2048, "n_mel": 40}),   # configs[4] sweep
    (44100, {}),                                                                        # Androids.conf:70 sampleRate
    (16000, {"frame_size": 0.032, "frame_step": 0.008, "n_mfcc": 19, "n_mel": 64, "cep_lifter": 0.0, "preemph": 0.0}),
    (16000, {"smooth_win": 0, "delta_win": 0}), (16000, {"smooth_win": 5, "delta_win": 0}), (16000, {"smooth_win": 1, "delta_win": 3}),
])
def test_lld_frames_and_functionals_match_the_numpy_restatement(ex, fs, params):
    """mshds_lld_extract (MFCC 1-12, RMS energy, ZCR per frame; mean / stddev per recording) against oracle/lld_oracle.py.
    Tolerance: 1e-9 relative / 1e-9 absolute (north_star: 1e-4 relative for spectral / MFCC statistics); ZCR exact."""
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200.synth import synth_clip
    rng = np.random.default_rng(4)
    clips = [synth_clip(600 + i, d, fs=fs).numpy() for i, d in enumerate([1.5, 0.73, 2.2])]
    clips += [np.zeros(0, np.int16), np.zeros(int(0.02 * fs), np.int16), np.zeros(int(0.3 * fs), np.int16),
              (rng.normal(scale=3000, size=int(0.5 * fs))).astype(np.int16)]
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    fun, frames, fo = ex.lld_extract(pcm, off, fs, want_frames=True, **params)
    wfun, wrows = lo.extract(pcm, off, float(fs), **params)
    D = wfun.shape[1] // 2
    assert list(np.diff(fo)) == [len(r) for r in wrows] and frames.shape == (int(fo[-1]), D)
    want = np.concatenate([r for r in wrows if len(r)])
    np.testing.assert_allclose(frames, want, rtol=1e-9, atol=1e-9)
    if params.get("smooth_win", 3) <= 1 and params.get("delta_win", 2) == 0:
        assert np.array_equal(frames[:, D - 1], want[:, D - 1])                  # raw zero-crossing rate: integer count / (nf - 1)
    assert np.array_equal(np.isnan(fun), np.isnan(wfun)) and np.isnan(fun[3]).all() and np.isnan(fun[4]).all()
    np.testing.assert_allclose(fun, wfun, rtol=1e-9, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("fs,params", [(16000, {}), (44100, {}), (16000, {"n_fft": 1024, "smooth_win": 0, "delta_win": 0})])
def test_lld_spectral_descriptors_and_the_twelve_functionals_match_the_numpy_restatement(ex, fs, params):
    """Second slice of the OpenSMILE path (descriptor_set = 1, functional_set = 1): cIntensity, 14 cSpectral descriptors and the
    twelve functionals of Androids.conf functL1 against oracle/lld_oracle.py.  Frames: 1e-8 relative, absolute tolerance scaled
    by the largest value of the column (band energies span 12 decades); roll-off bins and extreme positions are indices and
    must agree exactly; functionals: 1e-6 relative with the same column scaling (third / fourth moments of near-constant
    contours amplify the summation order)."""
    from oracle import lld_oracle as lo
    from robust_speech_analysis_framework_b200.lld_extractor import functional_names
    from robust_speech_analysis_framework_b200.synth import synth_clip
    rng = np.random.default_rng(5)
    clips = [synth_clip(620 + i, d, fs=fs).numpy() for i, d in enumerate([1.5, 0.73, 2.2])]
    clips += [np.zeros(0, np.int16), np.zeros(int(0.3 * fs), np.int16), (rng.normal(scale=3000, size=int(0.5 * fs))).astype(np.int16),
              synth_clip(630, 0.03, fs=fs).numpy()]                                     # the last one has a single frame
    pcm = np.concatenate(clips)
    off = np.cumsum([0] + [len(c) for c in clips]).astype(np.int64)
    kw = dict(descriptor_set=1, functional_set=1, **params)
    fun, frames, fo = ex.lld_extract(pcm, off, fs, want_frames=True, **kw)
    wfun, wrows = lo.extract(pcm, off, float(fs), **kw)
    names = functional_names(12, params.get("smooth_win", 3), params.get("delta_win", 2), 1, 1)
    assert len(names) == wfun.shape[1]
    _assert_lld_full_set_close(fun, frames, fo, wfun, wrows, raw=not params.get("smooth_win", 3) and not params.get("delta_win", 2))
    assert np.isnan(fun[3]).all()


def _assert_lld_full_set_close(fun, frames, fo, wfun, wrows, raw=False):
    """(functionals, frame rows, frame offsets) of the 720-column set against the wanted ones (see the test above for the
    tolerances); wrows: list of per-clip frame matrices."""
    W = wrows[0].shape[1]
    assert fun.shape == wfun.shape == (len(wrows), 12 * W)
    assert list(np.diff(fo)) == [len(r) for r in wrows] and frames.shape == (int(fo[-1]), W)
    want = np.concatenate([r for r in wrows if len(r)])
    scale = np.abs(want).max(axis=0) + 1e-300
    assert np.all(np.abs(frames - want) <= 1e-8 * np.abs(want) + 1e-9 * scale[None, :]), "frame rows"
    if raw:
        assert np.array_equal(frames[:, 18:22], want[:, 18:22])                         # raw roll-off points: bin frequencies
    assert np.array_equal(np.isnan(fun), np.isnan(wfun))
    for i in range(len(wrows)):
        if np.isnan(wfun[i]).all():
            continue
        g, w = fun[i].reshape(12, W), wfun[i].reshape(12, W)
        sc = np.abs(wrows[i]).max(axis=0) + 1e-300
        assert np.array_equal(g[3:5], w[3:5]), f"clip {i}: positions of the extremes"
        for k in (0, 1, 2, 5, 7, 9):                                                    # values in the contour's own unit
            assert np.all(np.abs(g[k] - w[k]) <= 1e-6 * np.abs(w[k]) + 1e-9 * sc), (i, k)
        T = len(wrows[i])
        assert np.all(np.abs(g[6] - w[6]) <= 1e-6 * np.abs(w[6]) + 1e-9 * sc / max(T, 1)), (i, "slope")
        assert np.all(np.abs(g[8] - w[8]) <= 1e-6 * np.abs(w[8]) + 1e-9 * sc * sc), (i, "regression error")
        np.testing.assert_allclose(g[10:12], w[10:12], rtol=1e-5, atol=1e-6, err_msg=f"clip {i}: skewness / kurtosis")


def test_lld_golden_vectors(ex):
    """tests/golden/lld_golden_v1.npz (outputs of the numpy restatement, frozen; generator: make_lld_golden.py): both descriptor
    sets of mshds_lld_extract against the committed numbers."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "lld_golden_v1.npz"))
    pcm, off = g["pcm"], g["offsets"]
    fun56, _, _ = ex.lld_extract(pcm, off, 16000)
    np.testing.assert_allclose(fun56, g["functionals_56"], rtol=1e-9, atol=1e-9, equal_nan=True)
    fun, frames, fo = ex.lld_extract(pcm, off, 16000, want_frames=True, descriptor_set=1, functional_set=1)
    counts = list(g["frame_counts"])
    starts = np.cumsum([0] + counts)
    wrows = [g["frames_720"][starts[i]:starts[i + 1]] for i in range(len(counts))]
    _assert_lld_full_set_close(fun, frames, fo, g["functionals_720"], wrows)


def test_lld_device_entry_and_bad_arguments(ex):
    import torch
    from robust_speech_analysis_framework_b200 import _lib
    pcm, off, _ = _batch([1.0, 1.3], start=80)
    host, _, _ = ex.lld_extract(pcm, off)
    d = torch.from_numpy(pcm).cuda()
    out = torch.empty((2, 56), dtype=torch.float64, device="cuda")
    ex.lld_extract_device(d.data_ptr(), off, out.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)
    for bad in ({"n_fft": 300}, {"n_fft": 256}, {"n_mel": 1}, {"n_mfcc": 26}, {"frame_step": 0.0}, {"mel_lo": 9000.0}, {"smooth_win": 4},
                {"delta_win": -1}, {"descriptor_set": 2}, {"functional_set": -1}):
        with pytest.raises(_lib.MshdsError):
            ex.lld_extract(pcm, off, **bad)


# ------------------------------------------------------------------------------------------------ frame-level contours (8f-4)
def test_frame_level_contours_match_the_oracle_objects(ex, orc):
    """mshds_extract_contours: the Pitch / Intensity / Harmonicity / Formant objects and the per-frame spectral moments behind
    the 25 columns, against the oracle's per-frame outputs (same tolerances as the stage-level test)."""
    pcm, off, clips = _batch([3.0, 2.50006, 0.02], start=90)            # the last clip is too short for any analysis
    want_feats, _ = orc.extract(pcm, off, 16000.0, nthreads=3)
    res = {k: ex.extract_contours(pcm, off, k) for k in ("f0", "intensity", "hnr", "formants", "moments")}
    for k, r in res.items():
        assert r["values"].shape[1] == {"f0": 2, "intensity": 1, "hnr": 1, "formants": 4, "moments": 4}[k]
        assert r["dt"] == 0.005 and len(r["frame_offsets"]) == 4 and r["frame_offsets"][3] == r["frame_offsets"][2]
        assert np.isnan(r["t1"][2])
        assert_features_close(r["features"], want_feats, f"features next to contour {k}")
    for c in range(2):
        x = orc.pcm_to_float(clips[c])
        fl, ce, _ = orc.pitch_values(x)
        sl = lambda r: r["values"][r["frame_offsets"][c]:r["frame_offsets"][c + 1]]
        ref = orc.pitch(x, 16000.0, 0, 0.005, fl, 3.0, 15, 0.03, 0.45, 0.01, 0.35, 0.14, ce)
        got = sl(res["f0"])
        assert len(got) == len(ref["freq"]) and abs(res["f0"]["t1"][c] - ref["x1"]) < 1e-12
        np.testing.assert_allclose(got[:, 0], ref["freq"], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(got[:, 1], ref["strength"], rtol=0, atol=1e-9)
        ic, ix1 = orc.intensity(x, 16000.0, fl, 0.005)
        np.testing.assert_allclose(sl(res["intensity"])[:, 0], ic, rtol=0, atol=1e-10)
        assert abs(res["intensity"]["t1"][c] - ix1) < 1e-12
        ph = orc.pitch(x, 16000.0, 2, 0.005, fl, 4.5, 15, 0.1, 0.0, 0.0, 0.0, 0.0, 8000.0)
        want_h = np.where(ph["freq"] == 0, -200.0, 10.0 * np.log10(np.maximum(ph["strength"], 1e-300) / (1.0 - ph["strength"])))
        np.testing.assert_allclose(sl(res["hnr"])[:, 0], want_h, rtol=1e-8, atol=1e-8)
        assert abs(np.mean(want_h[want_h != -200.0]) - want_feats[c, 9]) < 1e-8          # "Get mean" of the Harmonicity (:222)
        fo = orc.formants(x, 16000.0)
        gf = sl(res["formants"])
        assert len(gf) == len(fo["f"])
        for col, (arr, j) in enumerate(((fo["f"], 0), (fo["bw"], 0), (fo["f"], 1), (fo["bw"], 1))):
            w = np.where(fo["n"] > j, arr[:, j], np.nan)
            np.testing.assert_allclose(gf[:, col], w, rtol=1e-6, atol=1e-4, equal_nan=True)
        gm = sl(res["moments"])
        np.testing.assert_allclose(np.nanmean(gm, axis=0), want_feats[c, 21:25], rtol=1e-9)    # :371-374
    from robust_speech_analysis_framework_b200 import _lib
    with pytest.raises(KeyError):
        ex.extract_contours(pcm, off, "nope")


def test_float64_sample_entry(ex, orc, tmp_path):
    """MSHDS_PCM_FLOAT64: float64 samples (24/32-bit, multi-channel files) through the same kernels.  Samples that are exact
    int16 / 32768 values must give the int16 path's rows bit for bit when both take the frame-by-frame cross-correlation
    ("legacy_cc" = 1); by default int16 input takes the exact sliding sums, float64 input cannot (k_ccs.cu), so the harmonicity
    and formant columns then agree to rounding only.  A 24-bit-resolution signal is compared with the oracle."""
    import pandas as pd
    from src.mshds_extractor import extract_mshds_features
    pcm, off, clips = _batch([2.5, 3.1], start=95)
    b, sb = ex.extract_host_f64(pcm.astype(np.float64) / 32768.0, off)
    ex.set_option("legacy_cc", 1)
    try:
        a, sa = ex.extract_host(pcm, off)
    finally:
        ex.set_option("legacy_cc", 0)
    assert np.array_equal(a, b, equal_nan=True) and np.array_equal(sa, sb)
    a0, sa0 = ex.extract_host(pcm, off)
    assert_features_close(a0, b, "int16 (sliding-sum cross-correlation) vs float64 entry")
    assert np.array_equal(sa0, sb) and np.array_equal(a0[:, SPEECHRATE], b[:, SPEECHRATE])
    rng = np.random.default_rng(8)
    x = np.round((clips[0].astype(np.float64) / 32768.0 + rng.uniform(-0.5, 0.5, len(clips[0])) / 32768.0) * 8388608.0) / 8388608.0
    got, _ = ex.extract_host_f64(x, np.array([0, len(x)], np.int64))
    want, _ = orc.extract_f64(x, 16000.0)
    assert_features_close(got, want[None, :] if want.ndim == 1 else want, "24-bit resolution samples")
    # a stereo 16-bit file: Praat's channel mean has half-integer values; the DataFrame API routes it through the float64 entry
    st2 = np.stack([clips[1], np.clip(clips[1].astype(np.int32) + 1, -32768, 32767).astype(np.int16)], axis=1).ravel()
    path = str(tmp_path / "stereo.wav")
    with wave.open(path, "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(st2.astype("<i2").tobytes())
    df = extract_mshds_features(pd.DataFrame({"filepath": [path]}), verbose=False)
    mono = st2.reshape(-1, 2).astype(np.float64).mean(axis=1) / 32768.0
    wantm, _ = orc.extract_f64(mono, 16000.0)
    assert_features_close(df.iloc[:, 1:].to_numpy(dtype=np.float64), wantm[None, :] if wantm.ndim == 1 else wantm, "stereo file")
