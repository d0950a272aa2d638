// k_intensity.cu -- Sound_to_Intensity (fon/Sound_to_Intensity.cpp) and the Vector/Sampled queries on the contour.
//
// Serves mshds_extractor.py:41-52 (speech-rate intensity: 50 Hz, 16 ms; min / max "Parabolic", 0.99 quantile) and
// :198-202 (_extract_intensity: floor-dependent window, 5 ms; "Get mean 0 0 energy", min / max parabolic, ratio).
// One warp per frame: the Kaiser-20 weighted mean square of the (unweighted-)mean-removed span is two coalesced
// passes over <= 2049 int16 samples that stay L1-resident between neighbouring frames.
#include "internal.h"
#include "common.cuh"
#include "num.cuh"

__global__ void k_intensity_grid(Clips c, IntensityPass p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    long long nx = c.off[i + 1] - c.off[i];
    int k = p.class_dep ? c.cls[i] : 0;
    int nf = 0;
    double t1 = 0.0;
    bool ok = nx > 0 && short_term_analysis(nx, c.dx, c.x1[i], 6.4 / p.min_pitch[k], p.dt, &nf, &t1) != 0;
    p.nF[i] = ok ? nf : 0;
    p.t1[i] = t1;
}

__global__ void __launch_bounds__(256) k_intensity_frames(Clips c, IntensityPass p) {
    const int lane = threadIdx.x & 31;
    const int warpsPerBlock = blockDim.x >> 5;
    const int gw = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * warpsPerBlock;
    const int total = p.fstart[c.n];
    const double dx = c.dx;
    for (int f = gw; f < total; f += nwarps) {
        const int clip = find_segment(p.fstart, c.n, f);
        const double x1 = c.x1[clip];
        const int kcls = p.class_dep ? c.cls[clip] : 0;
        const int halfN = p.halfN[kcls];
        const double* __restrict__ win = p.win[kcls] + halfN;       // win[i], i in [-halfN, halfN]
        const long long base = c.off[clip], nx = c.off[clip + 1] - base;
        const SPtr pcm = c.pcm + base;
        const double midTime = p.t1[clip] + (double)(f - p.fstart[clip]) * p.dt;
        const long long midSample = x_to_nearest(x1, dx, midTime);
        long long leftSample = midSample - halfN, rightSample = midSample + halfN;
        if (leftSample < 1) leftSample = 1;
        if (rightSample > nx) rightSample = nx;
        double s = 0.0;
        for (long long i = leftSample + lane; i <= rightSample; i += 32) s += samp(pcm, i - 1);
        s = warp_sum(s);
        const double mean = s / (double)(rightSample - leftSample + 1);
        double sumxw = 0.0, sumw = 0.0;
        for (long long i = leftSample + lane; i <= rightSample; i += 32) {
            double a = samp(pcm, i - 1) - mean;
            double w = __ldg(win + (i - midSample));
            sumxw += a * a * w;
            sumw += w;
        }
        sumxw = warp_sum(sumxw);
        sumw = warp_sum(sumw);
        if (lane == 0) {
            double intensity = sumxw / sumw;
            intensity /= 4e-10;
            p.out[f] = intensity < 1e-30 ? -300.0 : 10.0 * log10(intensity);
        }
    }
}

void launch_intensity(const Clips& c, const IntensityPass& p, int max_frames_hint, cudaStream_t s) {
    k_intensity_grid<<<(c.n + 127) / 128, 128, 0, s>>>(c, p);
    launch_exclusive_scan(p.nF, p.fstart, c.n, s);
    int grid = (max_frames_hint + 7) / 8;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (grid < 1) grid = 1;
    k_intensity_frames<<<grid, 256, 0, s>>>(c, p);
}

// ------------------------------------------------------------------------------------------------ contour queries
__device__ __forceinline__ unsigned long long key_of(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double val_of(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}

// k-th smallest (1-based) of v[0..n) by MSB-first radix select; all threads of the CTA participate.
__device__ double block_select_kth(const double* __restrict__ v, int n, int k, int* hist /*[256]*/, unsigned long long* s_prefix,
                                   int* s_k) {
    unsigned long long prefix = 0, mask = 0;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            unsigned long long key = key_of(v[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 0xff)], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int kk = k, b = 0;
            for (; b < 256; b++) {
                if (kk <= hist[b]) break;
                kk -= hist[b];
            }
            if (b > 255) b = 255;
            *s_k = kk;
            *s_prefix = prefix | ((unsigned long long)b << shift);
        }
        __syncthreads();
        k = *s_k;
        prefix = *s_prefix;
        mask |= 0xffULL << shift;
        __syncthreads();
    }
    return val_of(prefix);
}

// Vector_getMaximum / getMinimum with parabolic interpolation over the whole domain, "Get mean 0 0 energy",
// "Get quantile 0 0 0.99".  stats[clip*4] = {min, max, q99, mean_energy_dB}.
__global__ void __launch_bounds__(256) k_contour_stats(Clips c, IntensityPass p, double* stats, int want_quantile) {
    __shared__ double red[32];
    __shared__ int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_k;
    const int clip = blockIdx.x;
    const int n = p.nF[clip];
    const double* __restrict__ y0 = p.out + p.fstart[clip];     // y0[i-1] = y[i]
    double mx = -CUDART_INF, mn = CUDART_INF, se = 0.0;
    if (n >= 1) {
        for (int i = 1 + threadIdx.x; i <= n; i += blockDim.x) {
            double yi = y0[i - 1];
            se += pow(10.0, 0.1 * yi);
            if (i == 1 || i == n) { mx = fmax(mx, yi); mn = fmin(mn, yi); }
            if (i >= 2 && i <= n - 1) {
                double yl = y0[i - 2], yr = y0[i];
                if (yi > yl && yi >= yr) {
                    double dy = 0.5 * (yr - yl), d2y = 2 * yi - yl - yr;
                    mx = fmax(mx, yi + 0.5 * dy * dy / d2y);
                }
                if (yi < yl && yi <= yr) {
                    double dy = 0.5 * (yr - yl), d2y = 2 * yi - yl - yr;
                    mn = fmin(mn, yi + 0.5 * dy * dy / d2y);
                }
            }
        }
    }
    mx = block_max(mx, red);
    mn = -block_max(-mn, red);
    se = block_sum(se, red);
    double q = DEVNAN;
    if (want_quantile && n >= 1) {
        double place = 0.99 * n + 0.5;
        int left = (int)floor(place);
        if (left < 1) q = block_select_kth(y0, n, 1, hist, &s_prefix, &s_k);
        else if (left >= n) q = block_select_kth(y0, n, n, hist, &s_prefix, &s_k);
        else {
            double a = block_select_kth(y0, n, left, hist, &s_prefix, &s_k);
            double b = block_select_kth(y0, n, left + 1, hist, &s_prefix, &s_k);
            q = (a == b) ? a : a + (place - left) * (b - a);
        }
    }
    if (threadIdx.x == 0) {
        stats[clip * 4 + 0] = n >= 1 ? mn : DEVNAN;
        stats[clip * 4 + 1] = n >= 1 ? mx : DEVNAN;
        stats[clip * 4 + 2] = q;
        stats[clip * 4 + 3] = n >= 1 ? 10.0 * log10(se / (double)n) : DEVNAN;
    }
}

void launch_contour_stats(const Clips& c, const IntensityPass& p, double* stats, int want_quantile, cudaStream_t s) {
    k_contour_stats<<<c.n, 256, 0, s>>>(c, p, stats, want_quantile);
}

// _extract_intensity (mshds_extractor.py:199-202)
__global__ void k_intensity_features(Clips c, IntensityPass p, const double* stats) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    if (p.nF[i] < 1) { atomicOr(&c.status[i], ST_INTENSITY); return; }
    double mn = stats[i * 4 + 0], mx = stats[i * 4 + 1];
    c.feat[(size_t)i * N_FEAT + 7] = stats[i * 4 + 3];
    c.feat[(size_t)i * N_FEAT + 8] = mn != 0.0 ? mx / mn : DEVNAN;
}
void launch_intensity_features(const Clips& c, const IntensityPass& p, const double* stats, cudaStream_t s) {
    k_intensity_features<<<(c.n + 127) / 128, 128, 0, s>>>(c, p, stats);
}
