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
#include "pitchq.cuh"

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
