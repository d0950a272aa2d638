// common.cuh -- shared device helpers for the B200-native MSHDS extractor.
//
// The whole library is compiled with -fmad=false so that every "a*b+c" written below rounds exactly like the
// float64 reference arithmetic (two roundings); hot inner loops that may fuse call fma() explicitly.  This matters
// where a floor()/round() of a time->sample-index conversion sits on an integer boundary (odd-length clips put
// every frame centre exactly on a sample), so index maths must be bit-identical to the CPU.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#define MSHDS_PI 3.14159265358979323846264338327950288
#define FULL_MASK 0xffffffffu
#define N_FEAT 25

// status bits, mirrored in include/mshds_b200.h
enum : uint32_t {
    ST_SPEECHRATE = 1u << 0, ST_PITCHRANGE_FALLBACK = 1u << 1, ST_PITCH = 1u << 2, ST_INTENSITY = 1u << 3,
    ST_HNR = 1u << 4, ST_LTAS = 1u << 5, ST_CPP = 1u << 6, ST_FORMANT = 1u << 7, ST_MOMENTS = 1u << 8,
    ST_FILE = 1u << 31
};

#define DEVNAN (__longlong_as_double(0x7ff8000000000000LL))

__host__ __device__ __forceinline__ bool is_undef(double x) { return !(x == x) || x > 1.7976931348623157e308 || x < -1.7976931348623157e308; }

// ---- Sampled index maths (fon/Sampled.h) for a sound that starts at 0: x1 = dx/2 -------------------------------
struct Grid {          // a Sampled: x1 + (i-1)*dx, i = 1..nx
    double x1, dx;
    int nx;
};
__host__ __device__ __forceinline__ double idx_to_x(double x1, double dx, double i) { return x1 + (i - 1.0) * dx; }
__host__ __device__ __forceinline__ double x_to_idx(double x1, double dx, double x) { return (x - x1) / dx + 1.0; }
__host__ __device__ __forceinline__ long long x_to_low(double x1, double dx, double x) { return (long long)floor(x_to_idx(x1, dx, x)); }
__host__ __device__ __forceinline__ long long x_to_high(double x1, double dx, double x) { return (long long)ceil(x_to_idx(x1, dx, x)); }
__host__ __device__ __forceinline__ long long iround_d(double x) { return (long long)floor(x + 0.5); }
__host__ __device__ __forceinline__ long long x_to_nearest(double x1, double dx, double x) { return iround_d(x_to_idx(x1, dx, x)); }

// Sampled_shortTermAnalysis: returns 0 if the sound is shorter than the window (Praat throws)
__host__ __device__ inline int short_term_analysis(long long nx, double dx, double x1, double windowDuration, double timeStep,
                                                   int* numberOfFrames, double* firstTime) {
    double myDuration = dx * (double)nx;
    if (windowDuration > myDuration) return 0;
    *numberOfFrames = (int)floor((myDuration - windowDuration) / timeStep) + 1;
    double ourMidTime = x1 - 0.5 * dx + 0.5 * myDuration;
    double thyDuration = (double)(*numberOfFrames) * timeStep;
    *firstTime = ourMidTime - 0.5 * thyDuration + 0.5 * timeStep;
    return 1;
}

// Sampled_getWindowSamples
__host__ __device__ inline long long get_window_samples(double x1, double dx, long long nx, double xmin, double xmax,
                                                        long long* ixmin, long long* ixmax) {
    double rixmin = 1.0 + ceil((xmin - x1) / dx);
    double rixmax = 1.0 + floor((xmax - x1) / dx);
    *ixmin = rixmin < 1.0 ? 1 : (long long)rixmin;
    *ixmax = rixmax > (double)nx ? nx : (long long)rixmax;
    if (*ixmin > *ixmax) return 0;
    return *ixmax - *ixmin + 1;
}

// ---- sample access: recordings stay int16 in HBM (2 B/sample); s = pcm/32768 exactly ---------------------------
// When the caller's audio is not at 16 kHz the front-end (Sound.resample(16000, 50), mshds_extractor.py:418-419) produces
// float64 samples on the device and the same kernels read those instead (p64 != nullptr; the branch is warp-uniform).
__device__ __forceinline__ double samp(const SPtr& s, long long i0 /*0-based*/) {
    return s.p64 ? __ldg(s.p64 + i0) : (double)__ldg(s.p16 + i0) * (1.0 / 32768.0);
}

// ---- warp / block reductions (fixed order => run-to-run deterministic) ------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
// block-wide sum; red must hold >= 32 doubles; every thread gets the result
__device__ __forceinline__ double block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = lane < nw ? red[lane] : 0.0;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ double block_max(double v, double* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = lane < nw ? red[lane] : -CUDART_INF;
    r = warp_max(r);
    return r;
}

// first index i in [0,n) with prefix[i+1] > key  (prefix is an exclusive scan with n+1 entries)
__device__ __forceinline__ int find_segment(const int* __restrict__ prefix, int n, int key) {
    int lo = 0, hi = n;   // invariant: prefix[lo] <= key < prefix[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= key) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int find_segment_ll(const long long* __restrict__ prefix, int n, long long key) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= key) lo = mid; else hi = mid;
    }
    return lo;
}

// speaker classes decided by _pitch_values (mshds_extractor.py:156-162)
#define CLS_MALE 0      // (60, 250)
#define CLS_FEMALE 1    // (100, 500)
#define CLS_FALLBACK 2  // (75, 500)
__host__ __device__ __forceinline__ double cls_floor(int c) { return c == 0 ? 60.0 : c == 1 ? 100.0 : 75.0; }
__host__ __device__ __forceinline__ double cls_ceiling(int c) { return c == 0 ? 250.0 : 500.0; }
