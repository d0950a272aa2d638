// k_lld.cu -- frame-level low-level descriptors of the reference's SECOND handcrafted extractor (OpenSMILE driven by
// Androids.conf; src/opensmile_extractor.py runs the external SMILExtract binary per file).  First slice of that path
// (SURVEY 8f-1): the cFramer -> cVectorPreemphasis -> cWindower -> cTransformFFT -> cFFTmagphase -> cMelspec -> cMfcc chain
// (Androids.conf:73-115), cEnergy rms (:117-123) and cMZcr zcr (:125-132), followed by mean / standard deviation over the
// frames of a recording (two of the cFunctionals of :functL1).  Definitions are spelled out in include/mshds_b200.h.
//
// One CTA per frame, 8 consecutive frames per turn (a 25 ms frame at a 10 ms hop shares 60 % of its samples with the next
// one: L1 hits; the batch is read from HBM once).  Pre-emphasis, Hamming window, the packed real FFT, the magnitude
// spectrum, the triangular mel bank, log, DCT-II and liftering all stay in shared memory; a frame leaves the SM as one row
// of n_mfcc + 2 doubles.
#include "internal.h"
#include "common.cuh"
#include "fft.cuh"

__global__ void k_lld_grid(int n, const long long* __restrict__ off, int nf, int ns, int* __restrict__ nF) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long nx = off[i + 1] - off[i];
    nF[i] = nx >= nf ? (int)((nx - nf) / ns) + 1 : 0;
}

__global__ void __launch_bounds__(128) k_lld_frames(LldPass p, const int16_t* __restrict__ pcm, const long long* __restrict__ off,
                                                     int n, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* a = (double2*)smem;                                  // M complex = n_fft reals (packed, swizzled)
    double* mag = (double*)(smem + sizeof(double2) * p.M);        // [M + 1] magnitude spectrum
    double* mel = mag + (p.M + 2);                                // [n_mel] log mel energies
    double* red = mel + ((p.n_mel + 1) & ~1);                     // [32]
    __shared__ int s_clip;
    const int total = p.fstart[n];
    const int D = p.n_mfcc + 2;
    for (int turn = blockIdx.x; turn * 8 < total; turn += gridDim.x)
    for (int f = turn * 8; f < total && f < turn * 8 + 8; f++) {
        __syncthreads();
        if (threadIdx.x == 0) s_clip = find_segment(p.fstart, n, f);
        __syncthreads();
        const int clip = s_clip;
        const int16_t* x = pcm + off[clip] + (long long)(f - p.fstart[clip]) * p.ns;      // frame samples x[0 .. nf)
        double* ar = (double*)a;
        double sq = 0.0, zc = 0.0;
        for (int j = threadIdx.x; j < p.n_fft; j += blockDim.x) {
            double v = 0.0;
            if (j < p.nf) {
                const double xj = (double)__ldg(x + j) * (1.0 / 32768.0);
                const double xm = j > 0 ? (double)__ldg(x + j - 1) * (1.0 / 32768.0) : 0.0;
                if (j > 0 && xj * xm < 0.0) zc += 1.0;                                  // cMZcr: sign changes of the raw frame
                const double pe = j > 0 ? xj - p.preemph * xm : xj * (1.0 - p.preemph);  // cVectorPreemphasis
                v = pe * __ldg(p.window + j);                                            // cWindower (Hamming)
                sq = fma(v, v, sq);
            }
            ar[SWZD(j)] = v;
        }
        sq = block_sum(sq, red);
        zc = block_sum(zc, red);
        __syncthreads();
        fft_dif<-1>(a, p.M, tw);
        // magnitude spectrum |X[k]|, k = 0..M, from the packed transform (bins sit bit-reversed)
        for (int k = threadIdx.x; k <= p.M / 2; k += blockDim.x) {
            if (k == 0) {
                const double2 z0 = a[SWZ(0)];
                mag[0] = fabs(z0.x + z0.y);
                mag[p.M] = fabs(z0.x - z0.y);
            } else {
                const int ik = bitrev(k, p.logM), imk = bitrev(p.M - k, p.logM);
                const double2 zk = a[SWZ(ik)], zmk = a[SWZ(imk)];
                const double2 wk = __ldg(tw + k * (TW_N / p.n_fft));
                double2 xk, xmk;
                real_bins_from_packed(zk, zmk, wk, &xk, &xmk);
                mag[k] = sqrt(xk.x * xk.x + xk.y * xk.y);
                mag[p.M - k] = sqrt(xmk.x * xmk.x + xmk.y * xmk.y);
            }
        }
        __syncthreads();
        // cMelspec (HTK-style triangles on the magnitude spectrum), then log
        for (int m = threadIdx.x; m < p.n_mel; m += blockDim.x) {
            const double c0 = __ldg(p.centres + m), c1 = __ldg(p.centres + m + 1), c2 = __ldg(p.centres + m + 2);
            double e = 0.0;
            for (int k = __ldg(p.klo + m); k <= __ldg(p.khi + m); k++) {
                const double mk = __ldg(p.melbin + k);
                const double up = (mk - c0) / (c1 - c0), down = (c2 - mk) / (c2 - c1);
                const double w = up < down ? up : down;
                if (w > 0.0) e = fma(w, mag[k], e);
            }
            mel[m] = log(e > p.log_floor ? e : p.log_floor);
        }
        __syncthreads();
        // cMfcc: DCT-II, coefficients 1..n_mfcc, sinusoidal liftering
        double* row = p.frames + (size_t)f * D;
        for (int i = threadIdx.x; i < p.n_mfcc; i += blockDim.x) {
            const int ci = i + 1;
            double c = 0.0;
            for (int m = 0; m < p.n_mel; m++) c = fma(mel[m], __ldg(p.dct + (size_t)i * p.n_mel + m), c);
            c *= p.dct_scale;
            if (p.lifter > 0.0) c *= 1.0 + 0.5 * p.lifter * sinpi((double)ci / p.lifter);
            row[i] = c;
        }
        if (threadIdx.x == 0) {
            row[p.n_mfcc] = sqrt(sq / (double)p.nf);                                   // cEnergy, rms of the windowed frame
            row[p.n_mfcc + 1] = p.nf > 1 ? zc / (double)(p.nf - 1) : 0.0;              // zero-crossing rate
        }
    }
}

// mean and population standard deviation of every descriptor over the frames of a clip (fixed-order reductions)
__global__ void __launch_bounds__(256) k_lld_functionals(LldPass p, int n, double* __restrict__ out) {
    __shared__ double red[32];
    const int clip = blockIdx.x;
    const int D = p.W;
    const int f0 = p.fstart[clip], nF = p.fstart[clip + 1] - f0;
    for (int d = 0; d < D; d++) {
        double s = 0.0;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) s += p.final[(size_t)(f0 + i) * D + d];
        s = block_sum(s, red);
        const double mean = nF > 0 ? s / (double)nF : DEVNAN;
        double v = 0.0;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) {
            const double e = p.final[(size_t)(f0 + i) * D + d] - mean;
            v = fma(e, e, v);
        }
        v = block_sum(v, red);
        if (threadIdx.x == 0) {
            out[(size_t)clip * 2 * D + d] = mean;
            out[(size_t)clip * 2 * D + D + d] = nF > 0 ? sqrt(v / (double)nF) : DEVNAN;
        }
        __syncthreads();
    }
}

// cContourSmoother (moving average over smooth_win frames, Androids.conf lld / lld2 / lld3) and cDeltaRegression (deltawin,
// Androids.conf delta1..3) on the frame rows of a clip; frames beyond the ends of the clip repeat the first / last frame.
// Output row: the D smoothed descriptors, then (delta_win > 0) their D regression deltas.
__global__ void __launch_bounds__(256) k_lld_post(LldPass p, int n) {
    const int D = p.n_mfcc + 2;
    const long long total = (long long)p.fstart[n] * D;
    const int hs = p.smooth_win > 1 ? p.smooth_win / 2 : 0;
    double dnorm = 0.0;
    for (int i = 1; i <= p.delta_win; i++) dnorm += 2.0 * i * i;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / D), d = (int)(t % D);
        const int clip = find_segment(p.fstart, n, f);
        const int f0 = p.fstart[clip], nF = p.fstart[clip + 1] - f0, k = f - f0;
        auto raw = [&](int kk) { kk = kk < 0 ? 0 : (kk >= nF ? nF - 1 : kk); return p.frames[(size_t)(f0 + kk) * D + d]; };
        auto sma = [&](int kk) {
            kk = kk < 0 ? 0 : (kk >= nF ? nF - 1 : kk);
            double a = 0.0;
            for (int u = -hs; u <= hs; u++) a += raw(kk + u);
            return a / (double)(2 * hs + 1);
        };
        double* row = p.final + (size_t)f * p.W;
        row[d] = sma(k);
        if (p.delta_win > 0) {
            double a = 0.0;
            for (int i = 1; i <= p.delta_win; i++) a += (double)i * (sma(k + i) - sma(k - i));
            row[D + d] = a / dnorm;
        }
    }
}

void launch_lld_post(const LldPass& p, int n, long long frames_hint, cudaStream_t s) {
    long long blocks = (frames_hint * (p.n_mfcc + 2) + 255) / 256;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    if (blocks < 1) blocks = 1;
    k_lld_post<<<(unsigned)blocks, 256, 0, s>>>(p, n);
}

void launch_lld_grid(int n, const long long* off, int nf, int ns, int* nF, int* fstart, cudaStream_t s) {
    k_lld_grid<<<(n + 127) / 128, 128, 0, s>>>(n, off, nf, ns, nF);
    launch_exclusive_scan(nF, fstart, n, s);
}
void launch_lld_frames(const LldPass& p, const int16_t* pcm, const long long* off, int n, const double2* tw, long long frames_hint,
                       cudaStream_t s) {
    const size_t smem = sizeof(double2) * p.M + sizeof(double) * (p.M + 2 + ((p.n_mel + 1) & ~1) + 32);
    cudaFuncSetAttribute(k_lld_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lld_frames, 128, smem);
    if (occ < 1) occ = 1;
    long long grid = (long long)sm_count() * occ;
    const long long nturn = (frames_hint + 7) / 8;
    if (grid > nturn) grid = nturn;
    if (grid < 1) grid = 1;
    k_lld_frames<<<(unsigned)grid, 128, smem, s>>>(p, pcm, off, n, tw);
}
void launch_lld_functionals(const LldPass& p, int n, double* out, cudaStream_t s) {
    if (n > 0) k_lld_functionals<<<n, 256, 0, s>>>(p, n, out);
}
