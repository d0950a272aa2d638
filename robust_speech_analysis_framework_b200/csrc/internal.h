// internal.h -- host-side declarations shared by the .cu translation units of libmshds_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Sample source of the chunk: int16 PCM as delivered, or float64 after the 16 kHz front-end resampler.
struct SPtr {
    const int16_t* p16;
    const double* p64;
    __host__ __device__ SPtr operator+(long long o) const { return SPtr{p16 ? p16 + o : nullptr, p64 ? p64 + o : nullptr}; }
    __host__ __device__ SPtr operator-(long long o) const { return SPtr{p16 ? p16 - o : nullptr, p64 ? p64 - o : nullptr}; }
};

// SM count of the current device (grids are sized in multiples of it), queried once per device
static inline int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;   // B200
        cached[dev] = n;
    }
    return cached[dev];
}

#define MAXCAND 15             // Pitch candidates kept per frame for the Viterbi passes (incl. the voiceless one)

// Per speaker-class configuration of one Sound_to_Pitch_any call (fon/Sound_to_Pitch.cpp set-up section).
struct PitchCfg {
    double floor_hz, ceiling, dt, ppw, dt_window, grid_window;
    double vt, octave_cost, sil, jump_cost, vuv_cost;
    int method;                // 0 = AC_HANNING, 2 = FCC_NORMAL
    int nsamp_period, halfnsamp_period, nsamp_window, halfnsamp_window, maximumLag, brent_ixmax;
    int nsampFFT, M, logM;     // AC only: real FFT size, packed complex size
    int maxn;                  // maxnCandidates after the floor(ceiling/floor) raise
    const double* window;      // [nsamp_window]   (AC)
    const double* windowR;     // [nsampFFT]       (AC) normalised window autocorrelation
};

// lags 0..len-1 of r are kept per frame for the refinement kernel; beyond maximumLag the FCC correlation is zero
static inline __host__ __device__ int stored_lags(const PitchCfg& g) {
    int len = g.brent_ixmax + 1;
    if (g.method >= 2 && g.maximumLag + 73 < len) len = g.maximumLag + 73;
    return len;
}

// One pitch analysis over the whole chunk (all clips).
struct PitchPass {
    PitchCfg cfg[3];           // indexed by speaker class; identical entries for class-independent passes
    int hnr_mode;              // 1: Sound_to_Harmonicity_cc (per-frame best strength only, no Viterbi)
    int* nF;                   // [n_clips]   frames per clip (0 when Praat would throw)
    double* t1;                // [n_clips]   time of first frame
    int* fstart;               // [n_clips+1] exclusive scan of nF
    // candidate scratch (shared between passes)
    double* cand_f;            // [frames*MAXCAND]
    double* cand_s;
    double* cand_score;        // local Viterbi score (delta)
    double* cand_lf;           // log2(f) of voiced candidates, -1 for voiceless
    uint8_t* ncand;            // [frames]
    uint8_t* psi;              // [frames*16] Viterbi back-pointers
    unsigned short* cand_imax; // [frames*MAXCAND] integer lag of the maximum behind each candidate
    double* inten;             // [frames] relative local peak (Pitch_Frame intensity)
    double* rbuf;              // [frames*rstride] correlation rows r[0..] kept for the refinement kernel
    int rstride;
    int* queue;                // refinement work list: frame*16 + candidate slot
    int* qcount;
    int* turn_counter;         // work counter of the frame kernel's persistent CTAs
    // dual mode: a second analysis that differs only in its voicing threshold shares frames, correlation and rbuf
    double dual_vt;
    double* dual_cand_f; double* dual_cand_s; unsigned short* dual_cand_imax; uint8_t* dual_ncand; double* dual_inten;
    int* dual_queue; int* dual_qcount;
    // harmonicity pass: every maximum of r is refined; items = frame<<32 | lag<<8 | sinc700 flag
    unsigned long long* queue64;
    unsigned long long* qcount64;
    unsigned long long q64_cap;
    unsigned long long* best_bits;   // [frames] bits of the largest refined strength (0 = none)
    // results
    double* sel_f;             // [frames] frequency of the chosen candidate (0 = voiceless)
    double* sel_s;             // [frames] strength of the chosen candidate / HNR: best r (NaN when voiceless)
};

struct Clips {
    int n;
    SPtr pcm;                  // packed samples of the chunk (int16, or float64 behind the resampling front-end)
    const long long* off;      // [n+1] sample offsets into pcm
    double fs, dx;
    const double* x1;          // [n] time of the first sample of each clip (dx/2 for a file read as is; shifted behind the
                               //     resampling front-end, where Praat centres the new samples in the old time domain)
    const double* xmax;        // [n] end of the clip's time domain (n*dx for a file read as is; the ORIGINAL duration otherwise)
    double* mean;              // [n] mean sample value
    double* gpeak;             // [n] max |s - mean|   (Sound_to_Pitch_any globalPeak)
    double* apeak;             // [n] max |s|          (Sound_Pitch_to_PointProcess_cc globalPeak)
    int* cls;                  // [n] speaker class (CLS_*)
    uint32_t* status;          // [n]
    double* feat;              // [n*25]
    long long total_samples;   // samples of the whole chunk behind pcm (bounds of the bulk copies)
    const double2* twb512;     // [32][16]  exp(-2 pi i j q / 512): pass twiddles of the warp-resident transforms (fftreg.cuh)
    const double2* twb1024;    // [32][32]  exp(-2 pi i j q / 1024)
    int legacy_fft;            // 1: CTA-per-frame shared-memory transforms (development switch "legacy_fft")
    int legacy_cc;             // development switch "legacy_cc": 1 = frame-by-frame cross-correlation, 2 = CTA block-ring kernel
};

// Glottal pulse sets (PointProcess) of one pitch pass
struct PulseSet {
    double cprime;             // raw slots per second of voiced stretch (ceiling / 0.8 with head-room)
    int* cap_start;            // [n+1] capacity offsets per clip (host-computed upper bounds)
    double* t;                 // final pulses of clip c: t[cap_start[c] .. +count[c])
    int* count;                // [n]
    int* st_count;             // [n]   voiced stretches per clip
    int* st_start;             // [n+1] exclusive scan
    int* st_ileft;             // [frames] first voiced frame (1-based) of stretch k of clip c at fstart[c]+k
    int* st_iright;            // last voiced frame
    double* raw_t;             // per-stretch raw points (left walk descending, then first + right walk)
    double* raw_thr;           // 0.8/f0 of left-walk points (for the addedRight rule)
    int* raw_nleft;            // [frames]
    int* raw_nright;
    double* raw_added_right;
    long long* raw_region;
    int* valid;                // [n] 1 when the pitch analysis behind this set exists (no Praat exception)
};
void launch_pulses(const Clips& c, const PitchPass& p, const PulseSet& ps, cudaStream_t s);

// "To Ltas (pitch-corrected)" accumulation (mshds_extractor.py:241-248)
struct LtasPass {
    int* part_count;           // [n]   groups of LTAS_PART pulses per clip
    int* part_start;           // [n+1]
    double* partial;           // [parts * 100] band energies (50) and counts (50)
    int* fail;                 // [n] set when Praat would throw
};
void launch_ltas(const Clips& c, const PulseSet& ps, const LtasPass& lt, double* ltas_bands /*[n*50]*/, cudaStream_t s);

// ---- launchers (each in its own .cu) ----------------------------------------------------------------------------
// frame-level descriptors of the OpenSMILE path (k_lld.cu)
struct LldPass {
    int nf, ns;                 // frame length / step in samples
    int n_fft, M, logM;         // transform size, M = n_fft / 2
    int n_mel, n_mfcc;
    int desc, fset, D0;         // descriptor_set, functional_set, raw descriptors per frame (n_mfcc + 2, + 16 with descriptor_set 1)
    double fs, win_sum;         // sampling frequency, sum of the Hamming window (cIntensity)
    double band_lo[2], band_hi[2], rolloff[4];   // cSpectral bands[] / rollOff[] (Androids.conf:261-266)
    double preemph, lifter, dct_scale, log_floor;
    const double* window;       // [nf] Hamming
    const double* melbin;       // [M + 1] mel value of every bin
    const double* centres;      // [n_mel + 2] band edges / centres on the mel axis
    const int* klo; const int* khi;   // [n_mel] first / last bin with a non-zero weight
    const double* dct;          // [n_mfcc * n_mel] cos(pi * i * (m + 0.5) / n_mel), i = 1..n_mfcc
    int* nF; int* fstart;       // [n], [n + 1]
    double* frames;             // [total_frames * (n_mfcc + 2)] raw descriptors
    int smooth_win, delta_win;  // contour smoothing window (frames, odd; <= 1 = none), delta regression half-window (0 = none)
    int W;                      // width of a final row: (n_mfcc + 2) * (delta_win > 0 ? 2 : 1)
    double* final;              // [total_frames * W]; == frames when neither smoothing nor deltas are requested
};
void launch_lld_grid(int n, const long long* off, int nf, int ns, int* nF, int* fstart, cudaStream_t s);
void launch_lld_frames(const LldPass& p, const int16_t* pcm, const long long* off, int n, const double2* tw, long long frames_hint,
                       cudaStream_t s);
void launch_lld_post(const LldPass& p, int n, long long frames_hint, cudaStream_t s);
void launch_lld_functionals(const LldPass& p, int n, double* out, cudaStream_t s);
void launch_session_agg(const double* feat, int n_cols, const int* row_start, const int* rows, int n_groups, double* mean_out,
                        double* std_out, cudaStream_t s);
void launch_finalize_status(const Clips& c, cudaStream_t s);
void launch_clip_stats(const Clips& c, long long max_clip_len, void* scratch /* n*16 bytes */, cudaStream_t s);
void launch_exclusive_scan(const int* counts, int* prefix, int n, cudaStream_t s);

void launch_pitch_grid(const Clips& c, const PitchPass& p, cudaStream_t s);
void launch_pitch_frames(const Clips& c, const PitchPass& p, const double2* tw, int max_frames_hint, cudaStream_t s);
void launch_pitch_refine(const Clips& c, const PitchPass& p, const double2* tw, int max_frames_hint, cudaStream_t s);
void launch_pitch_score(const Clips& c, const PitchPass& p, int max_frames_hint, cudaStream_t s);
bool launch_ac_frames_warp(const Clips& c, const PitchPass& p, const double2* tw, const double2* twb512, const double2* twb1024,
                           long long total_elems, int max_frames_hint, cudaStream_t s);      // k_acw.cu
bool launch_ac_candidates(const Clips& c, const PitchPass& p, const double2* tw, int max_frames_hint, cudaStream_t s);   // k_acw.cu
bool launch_cc_frames_shared(const Clips& c, const PitchPass& p, int max_frames_hint, cudaStream_t s);                // k_ccs.cu
bool launch_cc_frames_warp(const Clips& c, const PitchPass& p, int max_frames_hint, cudaStream_t s);                  // k_ccs.cu
void launch_pitch_viterbi(const Clips& c, const PitchPass& p, cudaStream_t s);
void launch_pitch_class(const Clips& c, const PitchPass& p, cudaStream_t s);                 // _pitch_values
void launch_pitch_stats(const Clips& c, const PitchPass& p, cudaStream_t s);                 // mean_F0, stdev semitones
void launch_hnr_mean(const Clips& c, const PitchPass& p, cudaStream_t s);


// Sound_to_Intensity pass
struct IntensityPass {
    int class_dep;             // 1: minimum pitch = speaker-class floor, 0: min_pitch[0]
    double min_pitch[3];
    double dt;
    int halfN[3];
    const double* win[3];      // [2*halfN+1] Kaiser-20 window per class
    int* nF; double* t1; int* fstart;
    double* out;               // [frames] dB
};
void launch_intensity(const Clips& c, const IntensityPass& p, int max_frames_hint, cudaStream_t s);
void launch_contour_stats(const Clips& c, const IntensityPass& p, double* stats /*[n*4] min,max,q99,mean_energy*/,
                          int want_quantile, cudaStream_t s);
void launch_intensity_features(const Clips& c, const IntensityPass& p, const double* stats, cudaStream_t s);

// Sound_to_Spectrogram + spectral moments pass
struct SpecPass {
    double physicalAnalysisWidth, timeStep, freqStep, y1, oneByBinWidth;
    int nsamp_window, halfnsamp_window, nsampFFT, M, logM, numberOfFreqs, binWidth_samples;
    const double* window;      // [nsamp_window] Gaussian (float32-rounded like Praat)
    int* nF; double* t1; int* fstart;
    double* mom;               // [frames*4]
    int* turn_counter;         // work counter of the warp-per-frame kernel's persistent CTAs
};
void launch_moments(const Clips& c, const SpecPass& p, const PitchPass& pp, const double2* tw, int max_frames_hint,
                    cudaStream_t s);

// _speechrate scratch (per clip: slots [fstart_sr[c] + 2c, +nF+2))
struct Ivl { double xmin, xmax; int sounding; int pad; };
struct SpeechRateScratch {
    Ivl* ivl;
    double* pk_t;
    double* pk_v;
    int* pk_i;
};
void launch_speechrate(const Clips& c, const IntensityPass& ip, const double* istats, const PitchPass& pp,
                       const SpeechRateScratch& sc, const double2* tw, cudaStream_t s);

// Sound_resample job: one sound (a whole clip or one voiced segment of it)
struct ResampleJob {
    long long clip_off;        // offset of the clip's first sample in the packed pcm
    long long nclip;           // clip length in samples
    long long ix1;             // 1-based clip index of source sample 1 (samples outside the clip are virtual zeros)
    long long nx;              // source samples
    double x1;                 // time of source sample 1 in the source sound's own time base
    long long zoff;            // complex work buffer offset (double2 units)
    long long filt_off;        // filtered float64 signal offset
    long long out_off;         // resampled signal offset
    long long nout;            // resampled samples
    double out_x1, out_dx;
    int table_id, pad;         // row block of the sinc coefficient table
};
void launch_resample_fft_group(const ResampleJob* d_jobs, const int* d_ids, int cnt, int logn, SPtr pcm, double2* zbuf,
                               double* filt, const double2* tw, double upfactor, cudaStream_t s, long long* launches, int mode = 0);
#define SINC_FIR_R 7        // outputs per thread of the polyphase FIR kernel (k_resample.cu FIR_R)
void launch_resample_copy(const ResampleJob* d_jobs, int njobs, long long max_nx, SPtr pcm, double* filt, cudaStream_t s, long long* launches);
void launch_sinc_resample(const ResampleJob* d_jobs, const long long* d_out_prefix, int njobs, long long total_out_hint,
                          const int* d_table_rep /*[ntables] representative job per table*/, int ntables,
                          const int* d_tile_prefix /*[njobs+1] FIR tiles per job*/, int total_tiles, const double* filt,
                          double* table, double* out, int P, int Q, int D, double dx_src, cudaStream_t s, long long* launches);

// Sound_to_Formant_burg over the resampled clips (job index == clip index)
struct FormantPass {
    const ResampleJob* jobs;
    double dt, dt_window, emphasis, nyquist;
    int nsamp_window;
    const double* window;      // [nsamp_window] Gaussian
    int* nF; double* t1; int* fstart;
    int* nform;                // [frames]
    double* freq;              // [frames*5]
    double* bw;                // [frames*5]
};
void launch_formants(const Clips& c, const FormantPass& p, int njobs, const double* sig, int max_frames_hint, cudaStream_t s);
void launch_formant_stats(const Clips& c, const FormantPass& p, const PulseSet& ps, cudaStream_t s);

// voiced segments of PointProcess_to_TextGrid_vuv (per clip, capacity-prefixed)
struct CppSegs {
    int* cap_start;            // [n+1]
    int* count;                // [n]
    int* fail;                 // [n]
    double* tmin; double* tmax;
    long long* ix1;
    int* nseg;
};
struct CepSeg {                // one cepstrogram (host-built from the segment list)
    double t1, windowDuration, dq;
    int nwin, nfft, logM, pad;
};
void launch_vuv_segments(const Clips& c, const PulseSet& ps, const CppSegs& sg, cudaStream_t s);
void launch_cepstrogram(const CepSeg* segs, const int* fprefix, int nsegs, const ResampleJob* jobs, const double* sig,
                        const double2* tw, double emphasis, double dt, double* cep, int nqmax, int total_frames,
                        const double* wtab, int wtab_n, cudaStream_t s, const double2* twb512 = nullptr, int* turn_counter = nullptr);
void launch_cpp_frames(const CepSeg* segs, const int* fprefix, int nsegs, const double* cep, int nqmax, int nTimeAvg,
                       double qAvgWindow, double* cpp_frame, int total_frames, cudaStream_t s);
void launch_cpp_reduce(const Clips& c, const CppSegs& sg, const int* seg_prefix, const int* fprefix, const double* cpp_frame,
                       cudaStream_t s);
