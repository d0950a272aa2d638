// k_lld.cu -- frame-level low-level descriptors of the reference's SECOND handcrafted extractor (OpenSMILE driven by
// Androids.conf; src/opensmile_extractor.py runs the external SMILExtract binary per file).  First slice of that path
// (SURVEY 8f-1): the cFramer -> cVectorPreemphasis -> cWindower -> cTransformFFT -> cFFTmagphase -> cMelspec -> cMfcc chain
// (Androids.conf:73-115), cEnergy rms (:117-123) and cMZcr zcr (:125-132); second slice (descriptor_set = 1): cIntensity
// intensity + loudness (:134-140) and 14 of the 16 cSpectral descriptors (:257-282: band energies, roll-off points, flux,
// centroid, entropy, variance, skewness, kurtosis, slope, flatness -- psychoacoustic sharpness and OpenSMILE's spectral
// harmonicity are not built), followed by functionals over the frames of a recording: mean / standard deviation, or
// (functional_set = 1) the twelve of :functL1 (Extremes, Regression, Moments).  Definitions are spelled out in
// include/mshds_b200.h.
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

// sums of N values over the CTA (fixed order); red holds >= 32 * N doubles; every thread gets the results
template <int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) red[i * 32 + w] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; i++) {
        double r = lane < nw ? red[i * 32 + lane] : 0.0;
        v[i] = warp_sum(r);
    }
}

#define LLD_NSPEC 16        // intensity, loudness, 2 band energies, 4 roll-off points, flux, centroid, entropy, variance, skewness, kurtosis, slope, flatness

// windowed frame -> packed FFT -> magnitude spectrum mag[0 .. M]; sums of the frame: sq = sum v^2, zc = sign changes of the
// raw frame, wi = sum w v^2 (cIntensity weights the windowed frame with a Hamming window of its own)
template <bool DESC>
__device__ __forceinline__ void lld_spectrum(const LldPass& p, const int16_t* __restrict__ x, double2* a, double* mag,
                                             const double2* __restrict__ tw, double& sq, double& zc, double& wi) {
    double* ar = (double*)a;
    sq = 0.0; zc = 0.0; wi = 0.0;
    for (int j = threadIdx.x; j < p.n_fft; j += blockDim.x) {
        double v = 0.0;
        if (j < p.nf) {
            const double xj = (double)__ldg(x + j) * (1.0 / 32768.0);
            const double xm = j > 0 ? (double)__ldg(x + j - 1) * (1.0 / 32768.0) : 0.0;
            if (j > 0 && xj * xm < 0.0) zc += 1.0;                                  // cMZcr: sign changes of the raw frame
            const double pe = j > 0 ? xj - p.preemph * xm : xj * (1.0 - p.preemph);  // cVectorPreemphasis
            const double w = __ldg(p.window + j);
            v = pe * w;                                                              // cWindower (Hamming)
            sq = fma(v, v, sq);
            if (DESC) wi = fma(w, v * v, wi);
        }
        ar[SWZD(j)] = v;
    }
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
}

template <bool DESC>
__global__ void __launch_bounds__(128) k_lld_frames(LldPass p, const int16_t* __restrict__ pcm, const long long* __restrict__ off,
                                                     int n, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* a = (double2*)smem;                                  // M complex = n_fft reals (packed, swizzled)
    double* mag = (double*)(smem + sizeof(double2) * p.M);        // [M + 1] magnitude spectrum
    double* mel = mag + (p.M + 2);                                // [n_mel] log mel energies
    double* red = mel + ((p.n_mel + 1) & ~1);                     // [32 * 8]
    double* magp = red + 32 * 8;                                  // [M + 1] previous frame's spectrum (descriptor_set = 1 only)
    __shared__ int s_clip;
    __shared__ double s_roll[4];
    const int total = p.fstart[n];
    const int D = p.D0;
    const int N = p.M + 1;
    const double df = p.fs / (double)p.n_fft;
    for (int turn = blockIdx.x; turn * 8 < total; turn += gridDim.x) {
    int have_prev = -2;                                           // frame whose spectrum sits in magp
    for (int f = turn * 8; f < total && f < turn * 8 + 8; f++) {
        __syncthreads();
        if (threadIdx.x == 0) s_clip = find_segment(p.fstart, n, f);
        __syncthreads();
        const int clip = s_clip;
        const int kf = f - p.fstart[clip];
        const int16_t* x = pcm + off[clip] + (long long)kf * p.ns;                      // frame samples x[0 .. nf)
        double sq, zc, wi;
        if (DESC && kf > 0 && have_prev != f - 1) {
            // first frame of a turn inside a recording: the spectral flux needs the spectrum of the frame before
            lld_spectrum<DESC>(p, x - p.ns, a, mag, tw, sq, zc, wi);
            for (int k = threadIdx.x; k < N; k += blockDim.x) magp[k] = mag[k];
            __syncthreads();
        }
        lld_spectrum<DESC>(p, x, a, mag, tw, sq, zc, wi);
        if (DESC) {
            double v3[3] = {sq, zc, wi};
            block_sum_n<3>(v3, red);
            sq = v3[0]; zc = v3[1]; wi = v3[2];
        } else {
            double v2[2] = {sq, zc};
            block_sum_n<2>(v2, red);
            sq = v2[0]; zc = v2[1];
        }
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
        if constexpr (DESC) {
            // ---- cSpectral on the power spectrum S[k] = |X[k]|^2, f_k = k fs / n_fft, k = 0 .. M (N = M + 1 bins)
            double v8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};      // sum S, sum f S, sum f, sum f^2 (by formula), flux, sum ln S, band0, band1
            for (int k = threadIdx.x; k < N; k += blockDim.x) {
                const double m = mag[k], S = m * m, fk = (double)k * df;
                v8[0] += S;
                v8[1] = fma(fk, S, v8[1]);
                if (kf > 0) { const double dm = m - magp[k]; v8[4] = fma(dm, dm, v8[4]); }
                v8[5] += log(S > 1e-100 ? S : 1e-100);
                if (fk >= p.band_lo[0] && fk <= p.band_hi[0]) v8[6] += S;
                if (fk >= p.band_lo[1] && fk <= p.band_hi[1]) v8[7] += S;
            }
            block_sum_n<8>(v8, red);
            const double sumS = v8[0];
            const double cen = sumS > 0.0 ? v8[1] / sumS : 0.0;
            double w4[4] = {0.0, 0.0, 0.0, 0.0};                           // central moments 2..4 (weighted by S), entropy
            for (int k = threadIdx.x; k < N; k += blockDim.x) {
                const double m = mag[k], S = m * m, d = (double)k * df - cen, d2 = d * d;
                w4[0] = fma(d2, S, w4[0]);
                w4[1] = fma(d2 * d, S, w4[1]);
                w4[2] = fma(d2 * d2, S, w4[2]);
                if (sumS > 0.0 && S > 0.0) { const double pk = S / sumS; w4[3] -= pk * log2(pk); }
            }
            block_sum_n<4>(w4, red);
            // roll-off points: first bin whose cumulative energy reaches the fraction (warp 0: contiguous chunks + warp scan)
            if (threadIdx.x < 32) {
                const int lane = threadIdx.x, per = (N + 31) / 32, k0 = lane * per, k1 = k0 + per < N ? k0 + per : N;
                double part = 0.0;
                for (int k = k0; k < k1; k++) part += mag[k] * mag[k];
                double inc = part;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(FULL_MASK, inc, o); if (lane >= o) inc += u; }
                const double before = inc - part, tot = __shfl_sync(FULL_MASK, inc, 31);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const double thr = p.rolloff[r] * tot;
                    int found = 0x7fffffff;
                    double cum = before;
                    for (int k = k0; k < k1; k++) { cum += mag[k] * mag[k]; if (cum >= thr) { found = k; break; } }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { const int u = __shfl_xor_sync(FULL_MASK, found, o); found = u < found ? u : found; }
                    if (lane == 0) s_roll[r] = (double)(found == 0x7fffffff ? N - 1 : found) * df;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                double* q = row + p.n_mfcc + 2;
                const double Im = (wi / p.win_sum) / 1e-6;                              // cIntensity: I / I0, I0 = 1e-6
                q[0] = Im;
                q[1] = pow(Im, 0.3);                                                    // loudness
                q[2] = v8[6]; q[3] = v8[7];
                q[4] = s_roll[0]; q[5] = s_roll[1]; q[6] = s_roll[2]; q[7] = s_roll[3];
                q[8] = kf > 0 ? sqrt(v8[4] / (double)N) : 0.0;                          // flux on magnitudes
                const double var = sumS > 0.0 ? w4[0] / sumS : 0.0, sd = sqrt(var);
                q[9] = cen;
                q[10] = w4[3];
                q[11] = var;
                q[12] = var > 0.0 ? (w4[1] / sumS) / (var * sd) : 0.0;
                q[13] = var > 0.0 ? (w4[2] / sumS) / (var * var) : 0.0;
                // slope of S over f: closed forms for sum f and sum f^2 over k = 0 .. M
                const double Nd = (double)N, Md = (double)p.M;
                const double sf = df * Md * Nd * 0.5, sff = df * df * Md * Nd * (2.0 * Md + 1.0) / 6.0;
                const double den = Nd * sff - sf * sf;
                q[14] = den != 0.0 ? (Nd * v8[1] - sf * sumS) / den : 0.0;
                q[15] = sumS > 0.0 ? exp(v8[5] / Nd) / (sumS / Nd) : 0.0;               // flatness: geometric / arithmetic mean
            }
            __syncthreads();
            for (int k = threadIdx.x; k < N; k += blockDim.x) magp[k] = mag[k];
            have_prev = f;
        }
    }
    }
}

// Functionals of every contour over the frames of a clip (fixed-order reductions).  functional_set 0: mean and population
// standard deviation; 1: the twelve of Androids.conf functL1 -- Extremes (max, min, range, maxPos, minPos, amean), Regression
// (linregc1, linregc2, linregerrQ), Moments (stddev, skewness, kurtosis); positions and the regression abscissa in frames.
__global__ void __launch_bounds__(256) k_lld_functionals(LldPass p, int n, double* __restrict__ out) {
    __shared__ double red[32 * 4];
    __shared__ double s_ext[2][8];
    __shared__ int s_pos[2][8];
    const int clip = blockIdx.x;
    const int D = p.W, NF = p.fset ? 12 : 2;
    const int f0 = p.fstart[clip], nF = p.fstart[clip + 1] - f0;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* o = out + (size_t)clip * NF * D;
    for (int d = 0; d < D; d++) {
        double a2[2] = {0.0, 0.0};                                  // sum y, sum t y
        double mx = -CUDART_INF, mn = CUDART_INF;
        int imx = 0x7fffffff, imn = 0x7fffffff;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) {
            const double y = p.final[(size_t)(f0 + i) * D + d];
            a2[0] += y;
            a2[1] = fma((double)i, y, a2[1]);
            if (y > mx) { mx = y; imx = i; }
            if (y < mn) { mn = y; imn = i; }
        }
        block_sum_n<2>(a2, red);
        if (p.fset) {
            // first position of the extremes: (value, index) pairs through the warps
#pragma unroll
            for (int s2 = 16; s2 > 0; s2 >>= 1) {
                const double ox = __shfl_xor_sync(FULL_MASK, mx, s2), on = __shfl_xor_sync(FULL_MASK, mn, s2);
                const int oix = __shfl_xor_sync(FULL_MASK, imx, s2), oin = __shfl_xor_sync(FULL_MASK, imn, s2);
                if (ox > mx || (ox == mx && oix < imx)) { mx = ox; imx = oix; }
                if (on < mn || (on == mn && oin < imn)) { mn = on; imn = oin; }
            }
            if (lane == 0) { s_ext[0][w] = mx; s_pos[0][w] = imx; s_ext[1][w] = mn; s_pos[1][w] = imn; }
            __syncthreads();
            mx = s_ext[0][0]; imx = s_pos[0][0]; mn = s_ext[1][0]; imn = s_pos[1][0];
            for (int k = 1; k < nw; k++) {
                if (s_ext[0][k] > mx || (s_ext[0][k] == mx && s_pos[0][k] < imx)) { mx = s_ext[0][k]; imx = s_pos[0][k]; }
                if (s_ext[1][k] < mn || (s_ext[1][k] == mn && s_pos[1][k] < imn)) { mn = s_ext[1][k]; imn = s_pos[1][k]; }
            }
        }
        const double T = (double)nF;
        const double mean = nF > 0 ? a2[0] / T : DEVNAN;
        // least-squares line y = m t + b over t = 0 .. T - 1
        const double tbar = 0.5 * (T - 1.0), stt = T * (T * T - 1.0) / 12.0;             // sum (t - tbar)^2
        const double m = nF > 1 ? (a2[1] - tbar * a2[0]) / stt : 0.0;
        const double b = mean - m * tbar;
        double c4[4] = {0.0, 0.0, 0.0, 0.0};                        // sum e^2, e^3, e^4, sum (y - line)^2
        for (int i = threadIdx.x; i < nF; i += blockDim.x) {
            const double y = p.final[(size_t)(f0 + i) * D + d];
            const double e = y - mean, e2 = e * e;
            c4[0] += e2;
            c4[1] = fma(e2, e, c4[1]);
            c4[2] = fma(e2, e2, c4[2]);
            const double r = y - fma(m, (double)i, b);
            c4[3] = fma(r, r, c4[3]);
        }
        block_sum_n<4>(c4, red);
        if (threadIdx.x == 0) {
            if (!p.fset) {
                o[d] = mean;
                o[D + d] = nF > 0 ? sqrt(c4[0] / T) : DEVNAN;
            } else if (nF < 1) {
                for (int k = 0; k < 12; k++) o[(size_t)k * D + d] = DEVNAN;
            } else {
                const double m2 = c4[0] / T, sd = sqrt(m2);
                o[0 * D + d] = mx; o[1 * D + d] = mn; o[2 * D + d] = mx - mn;
                o[3 * D + d] = (double)imx; o[4 * D + d] = (double)imn; o[5 * D + d] = mean;
                o[6 * D + d] = m; o[7 * D + d] = b; o[8 * D + d] = c4[3] / T;
                o[9 * D + d] = sd;
                // a contour that is constant up to rounding (stddev below 1e-12 of its largest value) has no shape
                const double amax = fmax(fabs(mx), fabs(mn));
                const bool flat = !(m2 > 1e-24 * amax * amax);
                o[10 * D + d] = flat ? 0.0 : (c4[1] / T) / (m2 * sd);
                o[11 * D + d] = flat ? 0.0 : (c4[2] / T) / (m2 * m2);
            }
        }
        __syncthreads();
    }
}

// cContourSmoother (moving average over smooth_win frames, Androids.conf lld / lld2 / lld3) and cDeltaRegression (deltawin,
// Androids.conf delta1..3) on the frame rows of a clip; frames beyond the ends of the clip repeat the first / last frame.
// Output row: the D smoothed descriptors, then (delta_win > 0) their D regression deltas.
__global__ void __launch_bounds__(256) k_lld_post(LldPass p, int n) {
    const int D = p.D0;
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
    long long blocks = (frames_hint * p.D0 + 255) / 256;
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
    const size_t smem = sizeof(double2) * p.M + sizeof(double) * (p.M + 2 + ((p.n_mel + 1) & ~1) + 32 * 8 + (p.desc ? p.M + 2 : 0));
    const void* kfn = p.desc ? (const void*)k_lld_frames<true> : (const void*)k_lld_frames<false>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, 128, smem);
    if (occ < 1) occ = 1;
    long long grid = (long long)sm_count() * occ;
    const long long nturn = (frames_hint + 7) / 8;
    if (grid > nturn) grid = nturn;
    if (grid < 1) grid = 1;
    if (p.desc) k_lld_frames<true><<<(unsigned)grid, 128, smem, s>>>(p, pcm, off, n, tw);
    else k_lld_frames<false><<<(unsigned)grid, 128, smem, s>>>(p, pcm, off, n, tw);
}
void launch_lld_functionals(const LldPass& p, int n, double* out, cudaStream_t s) {
    if (n > 0) k_lld_functionals<<<n, 256, 0, s>>>(p, n, out);
}
