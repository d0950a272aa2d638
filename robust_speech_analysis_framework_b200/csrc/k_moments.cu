// k_moments.cu -- Sound_to_Spectrogram (Gaussian, fon/Sound_and_Spectrogram.cpp) fused with Spectrogram_to_Spectrum
// and the power-weighted central moments of fon/Spectrum.cpp, masked by voicing.
//
// Serves _extract_Spectral_Moments (mshds_extractor.py:355-374).  The reference materialises a 320-bin spectrogram and
// walks it frame by frame from Python; here one CTA owns one spectrogram frame: window, packed real FFT-1024, 320 power
// bins and the four moments never leave shared memory, and only 4 doubles per voiced frame are written.
#include <cstdlib>
#include "internal.h"
#define NT_SPEC_DEFAULT 128
#include "common.cuh"
#include "fft.cuh"
#include "num.cuh"
#include "pitchq.cuh"
#include "fftwarp.cuh"
#include "stage.cuh"

__global__ void k_spec_grid(Clips c, SpecPass p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    long long nx = c.off[i + 1] - c.off[i];
    double duration = c.dx * (double)nx;
    int nf = 0;
    double t1 = 0.0;
    if (nx > 0 && !(p.physicalAnalysisWidth > duration)) {
        nf = 1 + (int)floor((duration - p.physicalAnalysisWidth) / p.timeStep);
        t1 = c.x1[i] + 0.5 * ((double)(nx - 1) * c.dx - (double)(nf - 1) * p.timeStep);
    }
    p.nF[i] = nf;
    p.t1[i] = t1;
}

__global__ void __launch_bounds__(256) k_spec_frames(Clips c, SpecPass p, PitchPass pp, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* a = (double2*)smem;                                  // M complex
    double* pw = (double*)(smem + sizeof(double2) * p.M);         // M+1 doubles
    double* red = pw + p.M + 8;
    __shared__ int s_clip, s_voiced;
    const int total = p.fstart[c.n];
    const double dx = c.dx;
    // 8 consecutive frames per turn: neighbouring frames share most of their samples (L1 hits)
    for (int turn = blockIdx.x; turn * 8 < total; turn += gridDim.x)
    for (int f = turn * 8; f < total && f < turn * 8 + 8; f++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int clip = find_segment(p.fstart, c.n, f);
            s_clip = clip;
            double t = p.t1[clip] + (double)(f - p.fstart[clip]) * p.timeStep;
            PitchView pv;
            pv.f = pp.sel_f + pp.fstart[clip]; pv.nx = pp.nF[clip]; pv.x1 = pp.t1[clip];
            pv.dx = pp.cfg[c.cls[clip]].dt; pv.ceiling = pp.cfg[c.cls[clip]].ceiling;
            pv.xmin = 0.0; pv.xmax = c.xmax[clip];
            s_voiced = pv.nx >= 1 && !is_undef(pitch_value_at(pv, t));
        }
        __syncthreads();
        if (!s_voiced) {                                           // mshds_extractor.py:364
            if (threadIdx.x < 4) p.mom[(size_t)f * 4 + threadIdx.x] = DEVNAN;
            continue;
        }
        const int clip = s_clip;
        const SPtr pcm = c.pcm + c.off[clip];
        const double t = p.t1[clip] + (double)(f - p.fstart[clip]) * p.timeStep;
        const long long leftSample = x_to_low(c.x1[clip], dx, t), rightSample = leftSample + 1;
        const long long startSample = rightSample - p.halfnsamp_window;
        double* ar = (double*)a;
        for (int m = threadIdx.x; m < p.nsampFFT; m += blockDim.x)
            ar[SWZD(m)] = m < p.nsamp_window ? samp(pcm, startSample + m - 1) * __ldg(p.window + m) : 0.0;
        __syncthreads();
        fft_dif<-1>(a, p.M, tw);
        packed_power_to_inverse_input(a, p.M, p.logM, tw, IdentityF(), pw);
        // power per band (binWidth_samples == 1 at 16 kHz; general case sums binWidth_samples FFT bins)
        double se = 0.0, sfe = 0.0;
        for (int ib = threadIdx.x; ib < p.numberOfFreqs; ib += blockDim.x) {
            double pwr = 0.0;
            for (int k = ib * p.binWidth_samples; k < (ib + 1) * p.binWidth_samples; k++) pwr += pw[k];
            pwr *= p.oneByBinWidth;
            double amp = sqrt(pwr);
            double e = amp * amp;
            ar[ib] = e;                                            // reuse the FFT buffer for the band energies
            se += e;
            sfe += (p.y1 + ib * p.freqStep) * e;
        }
        // NB: ar[] aliases a[]; packed_power_to_inverse_input ended with a barrier and pw is a separate array
        se = block_sum(se, red);
        sfe = block_sum(sfe, red);
        if (se == 0.0) {
            if (threadIdx.x < 4) p.mom[(size_t)f * 4 + threadIdx.x] = DEVNAN;
            continue;
        }
        const double fmean = sfe / se;
        double m2 = 0.0, m3 = 0.0, m4 = 0.0;
        for (int ib = threadIdx.x; ib < p.numberOfFreqs; ib += blockDim.x) {
            double e = ar[ib];
            double d = p.y1 + ib * p.freqStep - fmean;
            double d2 = d * d;
            m2 += d2 * e; m3 += d2 * d * e; m4 += d2 * d2 * e;
        }
        m2 = block_sum(m2, red); m3 = block_sum(m3, red); m4 = block_sum(m4, red);
        if (threadIdx.x == 0) {
            double mu2 = m2 / se, mu3 = m3 / se, mu4 = m4 / se;
            p.mom[(size_t)f * 4 + 0] = fmean;
            p.mom[(size_t)f * 4 + 1] = sqrt(mu2);
            p.mom[(size_t)f * 4 + 2] = mu2 != 0.0 ? mu3 / (mu2 * sqrt(mu2)) : DEVNAN;
            p.mom[(size_t)f * 4 + 3] = mu2 != 0.0 ? mu4 / (mu2 * mu2) - 3.0 : DEVNAN;
        }
    }
}


// ------------------------------------------------------------------------------------------------ warp-per-frame version
// Round 2: half a warp per spectrogram frame (1024-point real transform = 512 complex points = 16 lanes x 32 registers,
// fftreg.cuh / fftwarp.cuh), two frames per warp, the span of the 8 frames of a CTA turn staged by one bulk copy
// (stage.cuh).  The 320 band powers never exist outside registers: 20 per lane; the moments are 16-lane butterfly sums.
#define SPW_WARPS 4
#define SPW_NT (32 * SPW_WARPS)
#define SPW_TURN 8
#define SPW_XCH_BYTES (512 * 16 * 2)          // two frames of 512 complex doubles per warp
#define SPW_TW_BYTES (32 * 16 * 16)           // shared-memory copy of the [32][16] pass twiddles

struct SpwParams : StageParams {
    const double2* twb512;
    int* turn_counter;
};

__device__ __forceinline__ void spw_fetch(const Clips& c, const SpecPass& p, const SpwParams& A, int total, int nturn, int& cur_f,
                                          int& turn_end, StageSeg* sg, unsigned char* stage, unsigned long long* bar) {
    if (cur_f >= turn_end) {
        const int turn = atomicAdd(A.turn_counter, 1);
        if (turn >= nturn) { sg->n = 0; sg->tma_bytes = 0; return; }
        cur_f = turn * SPW_TURN;
        turn_end = cur_f + SPW_TURN < total ? cur_f + SPW_TURN : total;
    }
    const int clip = find_segment(p.fstart, c.n, cur_f);
    const int clip_end = p.fstart[clip + 1];
    const int f1 = turn_end < clip_end ? turn_end : clip_end;
    sg->f0 = cur_f; sg->n = f1 - cur_f; sg->clip = clip; sg->cls = c.cls[clip]; sg->k0 = cur_f - p.fstart[clip];
    const long long base = c.off[clip], nx = c.off[clip + 1] - base;
    const double x1 = c.x1[clip];
    const double tA = p.t1[clip] + (double)sg->k0 * p.timeStep, tB = p.t1[clip] + (double)(sg->k0 + sg->n - 1) * p.timeStep;
    long long sA = x_to_low(x1, c.dx, tA) + 1 - p.halfnsamp_window;                       // startSample of the first frame
    long long sB = x_to_low(x1, c.dx, tB) + 1 - p.halfnsamp_window + p.nsamp_window - 1;  // last sample of the last frame
    if (sA < 1) sA = 1;
    if (sB > nx) sB = nx;
    if (sB < sA) sB = sA - 1;
    stage_issue(A, sg, base, sA, (int)(sB - sA + 1), stage, bar);
    cur_f = f1;
}

__global__ void __launch_bounds__(SPW_NT, 2) k_spec_frames_w(const __grid_constant__ Clips c, const __grid_constant__ SpecPass p,
                                                             const __grid_constant__ PitchPass pp, const __grid_constant__ SpwParams A,
                                                             const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* xch_all = smem + SPW_TW_BYTES;                 // [pass twiddles][exchange x warps][2 stages]
    unsigned char* stage0 = xch_all + SPW_WARPS * SPW_XCH_BYTES;
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ StageSeg segs[2];
    constexpr int L = 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, j = lane & (L - 1), gidx = lane >> 4;
    const unsigned gmask = 0xffffu << (16 * gidx);
    const int total = p.fstart[c.n];
    const int nturn = (total + SPW_TURN - 1) / SPW_TURN;
    int cur_f = 0, turn_end = 0;
    if (tid == 0) mbar_init_pair(bars);
    fw_stage_twiddles<SPW_NT>((double2*)smem, A.twb512, 16);
    __syncthreads();
    if (tid == 0) spw_fetch(c, p, A, total, nturn, cur_f, turn_end, &segs[0], stage0, &bars[0]);
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    const double dx = c.dx;
    const int W = p.nsamp_window;
    const double2 wj = __ldg(tw + j * (TW_N / 1024));
    for (;;) {
        __syncthreads();
        const StageSeg& sg = segs[buf];
        if (sg.n == 0) break;
        unsigned char* st = stage0 + (size_t)buf * A.stage_bytes;
        if (tid == 0) spw_fetch(c, p, A, total, nturn, cur_f, turn_end, &segs[buf ^ 1], stage0 + (size_t)(buf ^ 1) * A.stage_bytes, &bars[buf ^ 1]);
        stage_complete<SPW_NT>(A, sg, st, &bars[buf], buf == 0 ? phase0 : phase1);
        const int clip = sg.clip;
        const int fi = 2 * warp + gidx;                            // SPW_TURN == 2 * SPW_WARPS: one round per segment
        const int f = sg.f0 + fi;
        bool active = fi < sg.n;
        const double t = p.t1[clip] + (double)(sg.k0 + fi) * p.timeStep;
        if (active) {                                              // mshds_extractor.py:364: frames with undefined pitch are skipped
            PitchView pv;
            pv.f = pp.sel_f + pp.fstart[clip]; pv.nx = pp.nF[clip]; pv.x1 = pp.t1[clip];
            pv.dx = pp.cfg[sg.cls].dt; pv.ceiling = pp.cfg[sg.cls].ceiling;
            pv.xmin = 0.0; pv.xmax = c.xmax[clip];
            const bool voiced = pv.nx >= 1 && !is_undef(pitch_value_at(pv, t));
            if (!voiced) {
                if (j < 4) p.mom[(size_t)f * 4 + j] = DEVNAN;
                active = false;
            }
        }
        if (__ballot_sync(FULL_MASK, active) != 0u) {
            double2* xch = (double2*)(xch_all + (size_t)warp * SPW_XCH_BYTES + (size_t)gidx * (SPW_XCH_BYTES / 2));
            const long long startSample = x_to_low(c.x1[clip], dx, t) + 1 - p.halfnsamp_window;
            const int ebase = (int)(startSample - sg.sA) + sg.shift;       // staged element of frame sample m = 0
            double2 a[32];
            fr_static_for<0, 32>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const int m = 2 * (j + L * k);
                double v0 = 0.0, v1 = 0.0;
                if (active && m < W) {                                     // W is even
                    const double2 w = __ldg((const double2*)(p.window + m));
                    v0 = staged(st, A.esz, ebase + m) * w.x;
                    v1 = staged(st, A.esz, ebase + m + 1) * w.y;
                }
                a[k] = make_double2(v0, v1);
            });
            fw_transform<L, -1>(a, xch, j, FwTwShared((const double2*)smem, j, L));
            double P[32];
            fw_powers<L, 20>(a, lane, j, wj, P);                           // bins j + 16 r, r < 20: the 320 bands of 15.625 Hz
            double se = 0.0, sfe = 0.0;
            fr_static_for<0, 20>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const int ib = j + L * r;
                const double amp = sqrt(P[r] * p.oneByBinWidth);           // Spectrogram_to_Spectrum: re = sqrt(power), im = 0
                const double e = amp * amp;                                // Spectrum moments with power 2: weight = re^2
                P[r] = e;
                se += e;
                sfe += (p.y1 + ib * p.freqStep) * e;
            });
            se = group_sum(se, gmask, L);
            sfe = group_sum(sfe, gmask, L);
            const double fmean = se != 0.0 ? sfe / se : 0.0;
            double m2 = 0.0, m3 = 0.0, m4 = 0.0;
            fr_static_for<0, 20>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const double d = p.y1 + (j + L * r) * p.freqStep - fmean;
                const double d2 = d * d;
                m2 += d2 * P[r]; m3 += d2 * d * P[r]; m4 += d2 * d2 * P[r];
            });
            m2 = group_sum(m2, gmask, L); m3 = group_sum(m3, gmask, L); m4 = group_sum(m4, gmask, L);
            if (active && j == 0) {
                if (se == 0.0) {
                    for (int k = 0; k < 4; k++) p.mom[(size_t)f * 4 + k] = DEVNAN;
                } else {
                    const double mu2 = m2 / se, mu3 = m3 / se, mu4 = m4 / se;
                    p.mom[(size_t)f * 4 + 0] = fmean;
                    p.mom[(size_t)f * 4 + 1] = sqrt(mu2);
                    p.mom[(size_t)f * 4 + 2] = mu2 != 0.0 ? mu3 / (mu2 * sqrt(mu2)) : DEVNAN;
                    p.mom[(size_t)f * 4 + 3] = mu2 != 0.0 ? mu4 / (mu2 * mu2) - 3.0 : DEVNAN;
                }
            }
        }
        buf ^= 1;
    }
}

static bool launch_spec_frames_warp(const Clips& c, const SpecPass& p, const PitchPass& pp, const double2* tw, int* turn_counter,
                                    int max_frames_hint, cudaStream_t s) {
    if (p.M != 512 || p.binWidth_samples != 1 || p.numberOfFreqs != 320 || (p.nsamp_window & 1) || p.nsamp_window > 1024 || !c.twb512) return false;
    SpwParams A;
    A.esz = c.pcm.p64 ? 8 : 2;
    A.pcm_bytes = c.pcm.p64 ? (const unsigned char*)c.pcm.p64 : (const unsigned char*)c.pcm.p16;
    A.total_elems = c.total_samples;
    const int hop = (int)ceil(p.timeStep / c.dx) + 1;
    A.stage_bytes = ((p.nsamp_window + (SPW_TURN - 1) * hop + 16) * A.esz + 32 + 127) & ~127;
    A.twb512 = c.twb512;
    A.turn_counter = turn_counter;
    const size_t smem = (size_t)SPW_TW_BYTES + (size_t)SPW_WARPS * SPW_XCH_BYTES + 2 * (size_t)A.stage_bytes;
    cudaMemsetAsync(turn_counter, 0, sizeof(int), s);
    cudaFuncSetAttribute(k_spec_frames_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_spec_frames_w, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spec_frames_w, SPW_NT, smem);
    if (occ < 1) occ = 1;
    int grid = sm_count() * occ;
    const int nturn = (max_frames_hint + SPW_TURN - 1) / SPW_TURN;
    if (max_frames_hint > 0 && grid > nturn) grid = nturn;
    if (grid < 1) grid = 1;
    k_spec_frames_w<<<grid, SPW_NT, smem, s>>>(c, p, pp, A, tw);
    return true;
}

// mean of each list over the frames that contributed (mshds_extractor.py:371-374)
__global__ void __launch_bounds__(256) k_spec_reduce(Clips c, SpecPass p, PitchPass pp) {
    __shared__ double red[32];
    const int clip = blockIdx.x;
    const int n = p.nF[clip], f0 = p.fstart[clip];
    for (int k = 0; k < 4; k++) {
        double s = 0.0, cnt = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double v = p.mom[(size_t)(f0 + i) * 4 + k];
            if (!is_undef(v)) { s += v; cnt += 1.0; }
        }
        s = block_sum(s, red); cnt = block_sum(cnt, red);
        if (threadIdx.x == 0) c.feat[(size_t)clip * N_FEAT + 21 + k] = (cnt > 0.0 && pp.nF[clip] >= 1) ? s / cnt : DEVNAN;
    }
    if (threadIdx.x == 0 && (n < 1 || pp.nF[clip] < 1)) atomicOr(&c.status[clip], ST_MOMENTS);
}

void launch_moments(const Clips& c, const SpecPass& p, const PitchPass& pp, const double2* tw, int max_frames_hint,
                    cudaStream_t s) {
    k_spec_grid<<<(c.n + 127) / 128, 128, 0, s>>>(c, p);
    launch_exclusive_scan(p.nF, p.fstart, c.n, s);
    if (!c.legacy_fft && p.turn_counter && launch_spec_frames_warp(c, p, pp, tw, p.turn_counter, max_frames_hint, s)) {
        k_spec_reduce<<<c.n, 256, 0, s>>>(c, p, pp);
        return;
    }
    size_t smem = sizeof(double2) * p.M + sizeof(double) * (p.M + 8 + 32);
    int grid = sm_count() * 8;
    if (max_frames_hint > 0 && grid > max_frames_hint) grid = max_frames_hint;
    if (grid < 1) grid = 1;
    cudaFuncSetAttribute(k_spec_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    static int nt = 0;
    if (!nt) { const char* e = getenv("MSHDS_NT_SPEC"); nt = e && atoi(e) == 256 ? 256 : (e && atoi(e) == 128 ? 128 : NT_SPEC_DEFAULT); }   // development switch
    if (nt == 128) { grid = sm_count() * 12; if (max_frames_hint > 0 && grid > max_frames_hint) grid = max_frames_hint; if (grid < 1) grid = 1; }
    k_spec_frames<<<grid, nt, smem, s>>>(c, p, pp, tw);
    k_spec_reduce<<<c.n, 256, 0, s>>>(c, p, pp);
}
