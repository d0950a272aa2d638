/*
 * mshds_b200.h -- C ABI of libmshds_b200.so, the B200-native (sm_100a) MSHDS acoustic feature extractor.
 *
 * The reference (ayushpradhan-dev/robust-speech-analysis-framework) has no FFI of its own: its hot path is the pure
 * Python function extract_mshds_features (src/mshds_extractor.py:379-459), whose arithmetic is Praat reached through
 * praat-parselmouth.  This header is the thin C layer BASELINE.json's north_star asks for; every entry point names the
 * reference interface it stands in for.  Plain pointers and sizes only, no torch types.  INTEGRATION.md shows the
 * ctypes binding and the drop-in src/mshds_extractor.py shim.
 *
 * Threading: one handle per host thread / process; a handle owns one CUDA device, its streams (the main one, which
 * mshds_set_stream can replace by the caller's, four side streams and a read-back stream) and all scratch memory;
 * no global mutable state.  Calls block until results are in the caller's buffers (mirrors the synchronous reference).
 */
#ifndef MSHDS_B200_H
#define MSHDS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSHDS_N_FEATURES 25

/* Column order of the output row == feature_names at src/mshds_extractor.py:397-404. */
enum mshds_feature {
    MSHDS_SPEAKING_RATE = 0, MSHDS_ARTICULATION_RATE, MSHDS_PHONATION_RATIO, MSHDS_PAUSE_RATE, MSHDS_MEAN_PAUSE_DURATION,
    MSHDS_MEAN_F0, MSHDS_STDEV_F0_SEMITONE, MSHDS_MEAN_DB, MSHDS_RANGE_RATIO_DB, MSHDS_HNR_DB,
    MSHDS_SPECTRAL_SLOPE, MSHDS_SPECTRAL_TILT, MSHDS_CPP,
    MSHDS_MEAN_F1, MSHDS_STD_F1, MSHDS_MEAN_B1, MSHDS_STD_B1, MSHDS_MEAN_F2, MSHDS_STD_F2, MSHDS_MEAN_B2, MSHDS_STD_B2,
    MSHDS_SPECTRAL_GRAVITY, MSHDS_SPECTRAL_STD_DEV, MSHDS_SPECTRAL_SKEWNESS, MSHDS_SPECTRAL_KURTOSIS
};

/* Per-clip status word.  A set bit means "this helper's try/except fired" in the reference
 * (src/mshds_extractor.py:124,161,182,204,224,250,300,337,375 and the whole-file handler :450-457): the columns of
 * that group are NaN.  Not an error of the call. */
#define MSHDS_ST_SPEECHRATE          (1u << 0)
#define MSHDS_ST_PITCHRANGE_FALLBACK (1u << 1)   /* _pitch_values returned (75, 500), :146,151,162 */
#define MSHDS_ST_PITCH               (1u << 2)
#define MSHDS_ST_INTENSITY           (1u << 3)
#define MSHDS_ST_HNR                 (1u << 4)
#define MSHDS_ST_LTAS                (1u << 5)
#define MSHDS_ST_CPP                 (1u << 6)
#define MSHDS_ST_FORMANT             (1u << 7)
#define MSHDS_ST_MOMENTS             (1u << 8)
#define MSHDS_ST_FILE                (1u << 31)  /* empty clip: whole row NaN (:450-457) */

/* flags of mshds_extract */
#define MSHDS_PCM_ON_DEVICE (1u << 0)   /* pcm is a device pointer on the handle's device */
#define MSHDS_OUT_ON_DEVICE (1u << 1)   /* features / status are device pointers */
#define MSHDS_PCM_FLOAT64   (1u << 2)   /* pcm points to float64 samples in [-1, 1) instead of int16 (24/32-bit and float files) */

/* error codes */
#define MSHDS_OK 0
#define MSHDS_ERR_ARG 1
#define MSHDS_ERR_CUDA 2
#define MSHDS_ERR_UNSUPPORTED 3

typedef struct mshds_handle mshds_handle;

/* Create / destroy an extractor bound to CUDA device `device`.  The reference keeps no state between calls
 * (src/mshds_extractor.py:379); the handle only caches window tables and scratch memory. */
int mshds_create(int device, mshds_handle** out);
void mshds_destroy(mshds_handle* h);

/* Issue all work of this handle on the caller's CUDA stream (cudaStream_t passed as void*), so that it is ordered after
 * whatever the caller queued there (e.g. the kernel that produced a device-resident pcm buffer).  NULL names the legacy
 * default stream -- what torch.cuda.current_stream().cuda_stream is outside a stream context -- NOT "no stream".
 * mshds_reset_stream goes back to the handle's private non-blocking stream (the state after mshds_create). */
int mshds_set_stream(mshds_handle* h, void* cuda_stream);
int mshds_reset_stream(mshds_handle* h);

/* Development / test switches; none selects a CPU path or changes a result.  Unknown names return MSHDS_ERR_ARG.
 *   "legacy_fft"     1: CTA-per-frame shared-memory FFT frame kernels (round 1) instead of the warp-per-frame register FFT with
 *                       TMA-staged sample spans; same results up to the rounding of another exact FFT order (A/B timing, tests)
 *   "legacy_cc"      1: frame-by-frame cross-correlation frames (round 1); 2: CTA block-ring kernel; 0 (default): one warp per
 *                       run of frames with exact sliding sums.  1 differs from 0 / 2 by rounding only, 0 and 2 are bit-identical
 *   "overlap"        0: issue the latency-bound per-clip kernels on the main stream instead of the side stream
 *   "nvtx"           1: emit one NVTX range per pipeline stage */
int mshds_set_option(mshds_handle* h, const char* name, long long value);

/* Upper bound, in samples, of the sub-batches the handle processes at once (scratch memory scales with it: ~170 B per sample).
 * Default: 2^27 samples (20 GB of scratch), grown automatically for batches of long recordings -- up to 128 recordings per
 * sub-batch, at most 2^29 samples and half of the free device memory -- because the per-recording sequential stages (path
 * finder, glottal-pulse walk) only fill the GPU through the number of recordings in flight.  Calling this fixes the size. */
int mshds_set_chunk_samples(mshds_handle* h, long long max_samples);

/* Human-readable description of the last non-zero return code of this handle. */
const char* mshds_last_error(const mshds_handle* h);

/*
 * The hot path: body of the per-file loop of extract_mshds_features (src/mshds_extractor.py:408-448), for a ragged
 * batch of mono recordings already decoded to 16-bit PCM (sample value = pcm / 32768, as parselmouth.Sound(path)
 * gives at :415).  Clip i is pcm[offsets[i] .. offsets[i+1]).  offsets is a HOST array of n_clips + 1 entries.
 * features: n_clips x 25 float64, row-major, column order as enum mshds_feature; status: n_clips words (may be NULL).
 * With MSHDS_PCM_FLOAT64 `pcm` points to float64 samples in [-1, 1) instead (what parselmouth.Sound holds for 24/32-bit,
 * float or multi-channel files after convert_to_mono, :415-417); offsets still count samples.
 * sample_rate is the rate of every clip of the call.  Anything but 16000 is first resampled to 16 kHz on the device exactly
 * as the reference does (:418-419 snd.resample(16000, 50): FFT brick-wall low-pass when down-sampling, sinc depth 50), and the
 * analyses then read the float64 result.  8000 Hz takes Praat's special case for an exact doubling (Sound_upsample: spectrum
 * tapered above 95 % of Nyquist, inverse transform of twice the length).  Returns MSHDS_OK or an error code; per-clip analysis
 * failures are NOT errors.
 */
int mshds_extract(mshds_handle* h, const int16_t* pcm, const int64_t* offsets, int n_clips, int sample_rate,
                  double* features, uint32_t* status, unsigned flags);

/*
 * Clip -> session aggregation, the step that follows the extractor in the reference pipeline: src/utils.py:36-56
 * aggregate_clip_features = groupby('unique_participant_id').agg(['mean', 'std']) over the clip rows.
 * features: n_rows x n_cols float64, row-major (NaN = missing, skipped like pandas does); row_group[i] in [0, n_groups) is
 * the session of row i, or negative for a row that belongs to none.  mean_out / std_out: n_groups x n_cols.  Rows are
 * visited in index order with pandas' own arithmetic (Kahan-compensated mean, Welford variance, ddof = 1: NaN for fewer
 * than two values), so the result is bit-identical to the reference's.  row_group is always a HOST array; with
 * MSHDS_AGG_ON_DEVICE features, mean_out and std_out are device pointers (e.g. the output of mshds_extract with
 * MSHDS_OUT_ON_DEVICE), otherwise host pointers.
 */
#define MSHDS_AGG_ON_DEVICE (1u << 0)
int mshds_aggregate_sessions(mshds_handle* h, const double* features, int n_rows, int n_cols, const int32_t* row_group,
                             int n_groups, double* mean_out, double* std_out, unsigned flags);

/*
 * Frame-level contours behind the 25 columns (the per-frame descriptors the dissertation names as the input a
 * sequence model would need, PDF p.32).  Runs the same pipeline as mshds_extract and copies ONE contour out per call:
 *   MSHDS_CONTOUR_F0         width 2: frequency in Hz (0 = unvoiced) and strength of the chosen candidate, the Pitch object
 *                            of _extract_pitch (mshds_extractor.py:178), one row per 5 ms frame
 *   MSHDS_CONTOUR_INTENSITY  width 1: dB, the Intensity object of :198
 *   MSHDS_CONTOUR_HNR        width 1: dB (-200 = voiceless), the Harmonicity object of :221
 *   MSHDS_CONTOUR_FORMANTS   width 4: F1, B1, F2, B2 in Hz (NaN = fewer formants in the frame), the Formant object of :319
 *   MSHDS_CONTOUR_MOMENTS    width 4: gravity, standard deviation, skewness, kurtosis of voiced spectrogram frames (NaN =
 *                            unvoiced, skipped at :364), the loop of :362-369
 * values: HOST buffer of capacity_rows x width doubles, clips back to back; frame_offsets (HOST, n_clips + 1): first row of
 * every clip; t1 (HOST, n_clips): time of a clip's first frame (NaN if it has none); *dt: frame step; frame k of clip i is at
 * t1[i] + k * dt.  A clip has at most floor(duration / 0.005) + 2 frames.  features (optional, HOST, n_clips x 25) receives
 * the feature rows as well.  flags: MSHDS_PCM_ON_DEVICE only.
 */
enum mshds_contour { MSHDS_CONTOUR_F0 = 0, MSHDS_CONTOUR_INTENSITY, MSHDS_CONTOUR_HNR, MSHDS_CONTOUR_FORMANTS, MSHDS_CONTOUR_MOMENTS };
int mshds_extract_contours(mshds_handle* h, const int16_t* pcm, const int64_t* offsets, int n_clips, int sample_rate, int contour,
                           double* values, size_t capacity_rows, int64_t* frame_offsets, double* t1, double* dt, int* width,
                           double* features, unsigned flags);

/*
 * Frame-level low-level descriptors + functionals: first slice of the reference's OTHER handcrafted extractor, the
 * OpenSMILE run of src/opensmile_extractor.py:9-103 with Androids.conf (the SMILExtract binary itself is
 * not available, so the component chain is restated here and in oracle/lld_oracle.py -- parity unpinned).  Covered:
 * cFramer -> cVectorPreemphasis -> cWindower -> cTransformFFT -> cFFTmagphase -> cMelspec -> cMfcc (Androids.conf:73-115),
 * cEnergy rms (:117-123), cMZcr zcr (:125-132), cContourSmoother and cDeltaRegression on those contours, and the mean /
 * standard deviation functionals over a recording.
 *
 * Definitions (fs = sample_rate, x = pcm / 32768):
 *   frames      nf = round(frame_size * fs) samples every ns = round(frame_step * fs); only complete frames:
 *               n_frames = nx >= nf ? (nx - nf) / ns + 1 : 0
 *   zcr         #{ j in 1..nf-1 : x[j] * x[j-1] < 0 } / (nf - 1), on the raw frame
 *   pre-emph    y[0] = (1 - k) * x[0],  y[j] = x[j] - k * x[j-1]
 *   window      Hamming  w[j] = 0.54 - 0.46 cos(2 pi j / (nf - 1))
 *   energy      sqrt( sum (y w)^2 / nf )
 *   spectrum    magnitude of the n_fft-point DFT of the zero-padded windowed frame, bins k = 0 .. n_fft/2
 *   mel bank    mel(f) = 1127 ln(1 + f / 700); n_mel + 2 points equally spaced on [mel(mel_lo), mel(min(mel_hi, fs/2))];
 *               band m = triangle over points m-1, m, m+1 evaluated at mel(k fs / n_fft); E_m = sum_k H_m(k) |X[k]|
 *   mfcc        c_i = sqrt(2 / n_mel) sum_m ln(max(E_m, 1e-10)) cos(pi i (m - 1/2) / n_mel), i = 1 .. n_mfcc,
 *               times 1 + (L / 2) sin(pi i / L) for cep_lifter L > 0
 *   smoothing   y_t = mean(x_{t-h} .. x_{t+h}), h = smooth_win / 2, frames beyond the ends of the recording repeat the end frame
 *   delta       d_t = sum_{i=1..delta_win} i (y_{t+i} - y_{t-i}) / (2 sum i^2), same end rule
 * Second slice, descriptor_set = 1 (Androids.conf:134-140 cIntensity, :257-282 cSpectral), 16 more descriptors per frame, in
 * this order; v = the pre-emphasised, windowed frame, |X[k]| its magnitude spectrum, S[k] = |X[k]|^2 (cSpectral squareInput),
 * f_k = k fs / n_fft, k = 0 .. n_fft/2 (N bins):
 *   intensity   (sum_j w[j] v[j]^2 / sum_j w[j]) / I0, I0 = 1e-6, w the Hamming window (cIntensity weights the frame once more)
 *   loudness    intensity^0.3
 *   fband       sum of S[k] over the bins with 250 <= f_k <= 650 Hz, and with 1000 <= f_k <= 4000 Hz (whole bins; OpenSMILE
 *               interpolates the two edge bins)
 *   rollOff     f_k of the first bin whose cumulative energy sum_{j<=k} S[j] reaches 25 / 50 / 75 / 90 % of sum S
 *   flux        sqrt( sum_k (|X_t[k]| - |X_t-1[k]|)^2 / N ), 0 for the first frame of a recording
 *   centroid    c = sum f_k S[k] / sum S;  variance = sum (f_k - c)^2 S[k] / sum S;  skewness, kurtosis = the third and fourth
 *               central moments over variance^1.5 and variance^2 (not excess);  entropy = - sum p_k log2 p_k, p_k = S[k] / sum S
 *   slope       least-squares slope of S[k] over f_k (per Hz);  flatness = exp(mean ln max(S[k], 1e-100)) / mean S[k]
 *   NOT built: psySharpness, spectralHarmonicity (:276,:278), the SHS pitch / Viterbi smoother (:142-213), jitter / shimmer
 *   (:231-248).  Open questions that cannot be settled without the SMILExtract binary are listed in DESIGN.md (float32
 *   internals, htkcompatible sample scaling, edge-bin interpolation of the bands, time normalisation of the functionals).
 * Row layout of a frame: the D (smoothed) descriptors mfcc[1..n_mfcc], energy, zcr [, the 16 above], then -- if
 * delta_win > 0 -- their D deltas: W = D or 2D values, D = n_mfcc + 2 [+ 16].
 * functionals: n_clips x NF x W, functional-major (NaN for a clip without a complete frame).  functional_set = 0: NF = 2,
 * amean then stddev (population).  functional_set = 1: NF = 12 in the order of Androids.conf functL1 (:349-366) --
 *   Extremes   max, min, range = max - min, maxPos, minPos (frame index of the first maximum / minimum), amean
 *   Regression linregc1 = slope m and linregc2 = offset b of the least-squares line y = m t + b over t = 0 .. T-1 (frames),
 *              linregerrQ = mean (y_t - (m t + b))^2
 *   Moments    stddev (population), skewness = m3 / m2^1.5, kurtosis = m4 / m2^2 (both 0 for a contour that is constant up to
 *              rounding: stddev <= 1e-12 of its largest absolute value)  frames_out (optional, may be NULL): all frame
 * rows, clips back to back; frame_offsets (optional HOST array, n_clips + 1) receives the first row of every clip.
 * flags: MSHDS_PCM_ON_DEVICE / MSHDS_OUT_ON_DEVICE as for mshds_extract (the latter covers functionals and frames_out).
 */
typedef struct mshds_lld_params {
    double frame_size;      /* s     0.025   Androids.conf:77  */
    double frame_step;      /* s     0.010   Androids.conf:78  */
    double preemph;         /*       0.97    Androids.conf:83  */
    int n_fft;              /*       0 = smallest power of two >= nf (cTransformFFT zero-pads); else a power of two in [nf, 8192] */
    int n_mel;              /*       26      cMelspec default nBands */
    double mel_lo, mel_hi;  /* Hz    20, 8000  Androids.conf:106-107 */
    int n_mfcc;             /*       12      Androids.conf:112-113 (coefficients 1..12) */
    double cep_lifter;      /*       22      cMfcc default */
    int smooth_win;         /* frames 3      cContourSmoother default smaWin (Androids.conf lld/lld2/lld3); <= 1: no smoothing */
    int delta_win;          /* frames 2      cDeltaRegression deltawin (Androids.conf delta1..3); 0: no delta columns */
    int descriptor_set;     /*       0       0: MFCC, energy, ZCR; 1: + cIntensity and cSpectral descriptors (16 more, see above) */
    int functional_set;     /*       0       0: amean, stddev; 1: the twelve functionals of Androids.conf functL1 (see above) */
} mshds_lld_params;
void mshds_lld_default_params(mshds_lld_params* p);
int mshds_lld_extract(mshds_handle* h, const int16_t* pcm, const int64_t* offsets, int n_clips, int sample_rate,
                      const mshds_lld_params* params, double* functionals, double* frames_out, int64_t* frame_offsets,
                      unsigned flags);

/* Number of kernel launches issued by this handle since creation (bench.py reports it as gpu_launches). */
long long mshds_launch_count(const mshds_handle* h);

/*
 * Per-stage device timing (CUDA events on the handle's stream, accumulated over calls while enabled).  The report is
 * text: one "stage<TAB>milliseconds<TAB>spans" line per stage in first-use order.  bench.py uses it for the roofline
 * figure of the dominant kernel; the reference has no counterpart (it only has a tqdm bar, mshds_extractor.py:406).
 */
int mshds_profile_enable(mshds_handle* h, int on);
/* Measured float64 FMA issue peak of the handle's device in TFLOP/s (2 FLOP per DFMA; best of five launches of a
 * register-only kernel with 8 independent chains per thread): the denominator of the compute-side roofline bench.py reports,
 * since MEASURED_PEAKS.json only holds HBM and bf16 tensor figures and this path is float64 on the vector pipe. */
int mshds_fp64_peak(mshds_handle* h, double* tflops);
int mshds_profile_report(mshds_handle* h, char* buf, size_t cap);

/*
 * Stage-level read-back for parity tests: after mshds_extract, copy one named intermediate of clip `clip` of the LAST
 * processed chunk into a host buffer.  Names: "pitch_wide_f", "pitch_main_f", "pitch_main_s", "pitch_cpp_f",
 * "pitch_ltas_f", "pitch_cc_f", "pitch_sr_f", "hnr_r", "intensity_main", "intensity_sr", "pulses_cpp", "pulses_fmt",
 * "pulses_ltas", "ltas_bands", "formant_f", "formant_b", "formant_n", "resampled10k", "class", "moments".
 * Returns the number of elements available through *n_out (copies at most cap elements); element type is float64
 * except "class"/"formant_n" (int32).
 */
int mshds_debug_fetch(mshds_handle* h, const char* name, int clip, void* host_buf, size_t cap_elems, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
