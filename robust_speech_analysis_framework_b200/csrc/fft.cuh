// fft.cuh -- shared-memory float64 FFT building blocks (hand-written, no cuFFT).
//
// Complex radix-2 in place: DIF (natural in -> bit-reversed out) for the forward transform and DIT (bit-reversed in
// -> natural out) for the inverse, so no explicit permutation pass is ever run.  Real transforms use the packed
// trick (N real samples as N/2 complex) with the untangle / power / retangle step done pairwise in bit-reversed
// position.  Twiddles come from one global table tw[j] = exp(-2*pi*i*j/TW_N), j < TW_N/2 (L1/L2 resident).
#pragma once
#include "common.cuh"

#define TW_N 8192          // table covers complex FFT sizes up to 8192

// Shared-memory swizzle of the single-column transforms: complex element i lives at i ^ ((i >> 2) & 7).  The fused passes
// touch elements 4t, 4t+1.. (late DIF / early DIT passes) or t, t+q.. (early passes); with 16-byte elements an unswizzled
// layout puts a quarter-warp on 2-4 of the 8 bank groups, the XOR spreads every such pattern over all 8.
#define SWZ(i) ((i) ^ (((i) >> 2) & 7))
#define SWZD(m) ((SWZ((m) >> 1) << 1) | ((m) & 1))      // same for a flat float64 view of the buffer

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ int bitrev(int k, int logn) { return (int)(__brev((unsigned)k) >> (32 - logn)); }

// tw index for exp(SIGN*2*pi*i*pos/len), len a power of two <= TW_N, pos < len/2
template <int SIGN>
__device__ __forceinline__ double2 twiddle(const double2* __restrict__ tw, int pos, int len) {
    double2 w = __ldg(tw + (pos << (13 - (__ffs(len) - 1))));          // pos * (TW_N / len), TW_N = 2^13
    if (SIGN > 0) w.y = -w.y;
    return w;
}

// n complex points, cols interleaved columns (element (r,c) at a[r*cols+c]); all threads of the block participate.
template <int SIGN>
__device__ __forceinline__ void fft_dif_cols(double2* a, int n, int cols, const double2* __restrict__ tw) {
    for (int half = n >> 1; half >= 1; half >>= 1) {
        int work = (n >> 1) * cols;
        for (int j = threadIdx.x; j < work; j += blockDim.x) {
            int c = j % cols, b = j / cols;
            int pos = b & (half - 1), grp = b / half;
            int i0 = (grp * 2 * half + pos) * cols + c, i1 = i0 + half * cols;
            double2 u = a[i0], v = a[i1];
            a[i0] = make_double2(u.x + v.x, u.y + v.y);
            double2 d = make_double2(u.x - v.x, u.y - v.y);
            a[i1] = half > 1 ? cmul(d, twiddle<SIGN>(tw, pos, 2 * half)) : d;
        }
        __syncthreads();
    }
}
template <int SIGN>
__device__ __forceinline__ void fft_dit_cols(double2* a, int n, int cols, const double2* __restrict__ tw) {
    for (int half = 1; half < n; half <<= 1) {
        int work = (n >> 1) * cols;
        for (int j = threadIdx.x; j < work; j += blockDim.x) {
            int c = j % cols, b = j / cols;
            int pos = b & (half - 1), grp = b / half;
            int i0 = (grp * 2 * half + pos) * cols + c, i1 = i0 + half * cols;
            double2 u = a[i0], v = a[i1];
            if (half > 1) v = cmul(v, twiddle<SIGN>(tw, pos, 2 * half));
            a[i0] = make_double2(u.x + v.x, u.y + v.y);
            a[i1] = make_double2(u.x - v.x, u.y - v.y);
        }
        __syncthreads();
    }
}

// Single-column versions (the per-frame transforms).  Two radix-2 stages are fused per pass (4 points in registers, the
// second pair's first-stage twiddle is the first one rotated by a quarter turn), which halves the shared-memory round
// trips, the barriers and the twiddle loads while keeping the plain radix-2 bit-reversed ordering.
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
template <int SIGN>
__device__ __forceinline__ double2 quarter_turn(double2 a) {      // a * exp(SIGN * i * pi / 2)
    return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// a * exp(SIGN * i * pi / 4) and a * exp(SIGN * 3 i pi / 4)
template <int SIGN>
__device__ __forceinline__ double2 eighth_turn(double2 a) {
    const double r = 0.70710678118654752440;
    return SIGN > 0 ? make_double2((a.x - a.y) * r, (a.x + a.y) * r) : make_double2((a.x + a.y) * r, (a.y - a.x) * r);
}
template <int SIGN>
__device__ __forceinline__ double2 three_eighth_turn(double2 a) {
    const double r = 0.70710678118654752440;
    return SIGN > 0 ? make_double2(-(a.x + a.y) * r, (a.x - a.y) * r) : make_double2((a.y - a.x) * r, -(a.x + a.y) * r);
}

// Three radix-2 DIF stages (spans half, half/2, half/4) fused: 8 points in registers, one shared-memory round trip and one
// barrier instead of two (radix-2x2 + radix-2).  Same stage sequence, same bit-reversed ordering as the plain radix-2 form.
template <int SIGN>
__device__ __forceinline__ void fft_dif_pass8(double2* a, int n, int half, const double2* __restrict__ tw) {
    const int q = half >> 2;
    const int lq = __ffs(q) - 1;
    for (int t = threadIdx.x; t < (n >> 3); t += blockDim.x) {
        const int pos = t & (q - 1);
        const int i0 = ((t >> lq) << (lq + 3)) + pos;
        double2 x0 = a[SWZ(i0)], x1 = a[SWZ(i0 + q)], x2 = a[SWZ(i0 + 2 * q)], x3 = a[SWZ(i0 + 3 * q)];
        double2 x4 = a[SWZ(i0 + 4 * q)], x5 = a[SWZ(i0 + 5 * q)], x6 = a[SWZ(i0 + 6 * q)], x7 = a[SWZ(i0 + 7 * q)];
        const double2 w8 = twiddle<SIGN>(tw, pos, 2 * half);
        const double2 w4 = cmul(w8, w8);
        const double2 w2 = cmul(w4, w4);
        // span half: (j, j + 4), twiddle w8 * exp(SIGN 2 pi i j / 8)
        double2 s0 = cadd(x0, x4), s1 = cadd(x1, x5), s2 = cadd(x2, x6), s3 = cadd(x3, x7);
        double2 d0 = cmul(csub(x0, x4), w8);
        double2 d1 = cmul(eighth_turn<SIGN>(csub(x1, x5)), w8);
        double2 d2 = cmul(quarter_turn<SIGN>(csub(x2, x6)), w8);
        double2 d3 = cmul(three_eighth_turn<SIGN>(csub(x3, x7)), w8);
        // span half/2: (j, j + 2) inside each half, twiddle w4 * exp(SIGN 2 pi i (j & 1) / 4)
        double2 a0 = cadd(s0, s2), a1 = cadd(s1, s3);
        double2 b0 = cmul(csub(s0, s2), w4), b1 = cmul(quarter_turn<SIGN>(csub(s1, s3)), w4);
        double2 c0 = cadd(d0, d2), c1 = cadd(d1, d3);
        double2 e0 = cmul(csub(d0, d2), w4), e1 = cmul(quarter_turn<SIGN>(csub(d1, d3)), w4);
        // span half/4: neighbours, twiddle w2
        a[SWZ(i0)] = cadd(a0, a1);
        a[SWZ(i0 + q)] = cmul(csub(a0, a1), w2);
        a[SWZ(i0 + 2 * q)] = cadd(b0, b1);
        a[SWZ(i0 + 3 * q)] = cmul(csub(b0, b1), w2);
        a[SWZ(i0 + 4 * q)] = cadd(c0, c1);
        a[SWZ(i0 + 5 * q)] = cmul(csub(c0, c1), w2);
        a[SWZ(i0 + 6 * q)] = cadd(e0, e1);
        a[SWZ(i0 + 7 * q)] = cmul(csub(e0, e1), w2);
    }
    __syncthreads();
}
// the DIT mirror: spans h, 2h, 4h
template <int SIGN>
__device__ __forceinline__ void fft_dit_pass8(double2* a, int n, int h, const double2* __restrict__ tw) {
    const int lh = __ffs(h) - 1;
    for (int t = threadIdx.x; t < (n >> 3); t += blockDim.x) {
        const int pos = t & (h - 1);
        const int i0 = ((t >> lh) << (lh + 3)) + pos;
        double2 x0 = a[SWZ(i0)], x1 = a[SWZ(i0 + h)], x2 = a[SWZ(i0 + 2 * h)], x3 = a[SWZ(i0 + 3 * h)];
        double2 x4 = a[SWZ(i0 + 4 * h)], x5 = a[SWZ(i0 + 5 * h)], x6 = a[SWZ(i0 + 6 * h)], x7 = a[SWZ(i0 + 7 * h)];
        const double2 w8 = twiddle<SIGN>(tw, pos, 8 * h);
        const double2 w4 = cmul(w8, w8);
        const double2 w2 = cmul(w4, w4);
        // span h: neighbours, twiddle w2
        x1 = cmul(x1, w2); x3 = cmul(x3, w2); x5 = cmul(x5, w2); x7 = cmul(x7, w2);
        double2 y0 = cadd(x0, x1), y1 = csub(x0, x1), y2 = cadd(x2, x3), y3 = csub(x2, x3);
        double2 y4 = cadd(x4, x5), y5 = csub(x4, x5), y6 = cadd(x6, x7), y7 = csub(x6, x7);
        // span 2h: (j, j + 2), twiddle w4 * exp(SIGN 2 pi i (j & 1) / 4)
        y2 = cmul(y2, w4); y3 = cmul(quarter_turn<SIGN>(y3), w4); y6 = cmul(y6, w4); y7 = cmul(quarter_turn<SIGN>(y7), w4);
        double2 z0 = cadd(y0, y2), z2 = csub(y0, y2), z1 = cadd(y1, y3), z3 = csub(y1, y3);
        double2 z4 = cadd(y4, y6), z6 = csub(y4, y6), z5 = cadd(y5, y7), z7 = csub(y5, y7);
        // span 4h: (j, j + 4), twiddle w8 * exp(SIGN 2 pi i j / 8)
        z4 = cmul(z4, w8);
        z5 = cmul(eighth_turn<SIGN>(z5), w8);
        z6 = cmul(quarter_turn<SIGN>(z6), w8);
        z7 = cmul(three_eighth_turn<SIGN>(z7), w8);
        a[SWZ(i0)] = cadd(z0, z4);
        a[SWZ(i0 + 4 * h)] = csub(z0, z4);
        a[SWZ(i0 + h)] = cadd(z1, z5);
        a[SWZ(i0 + 5 * h)] = csub(z1, z5);
        a[SWZ(i0 + 2 * h)] = cadd(z2, z6);
        a[SWZ(i0 + 6 * h)] = csub(z2, z6);
        a[SWZ(i0 + 3 * h)] = cadd(z3, z7);
        a[SWZ(i0 + 7 * h)] = csub(z3, z7);
    }
    __syncthreads();
}

#ifndef FFT_RADIX8
#define FFT_RADIX8 1
#endif

template <int SIGN>
__device__ __forceinline__ void fft_dif(double2* a, int n, const double2* __restrict__ tw) {
    int half = n >> 1;
#if FFT_RADIX8
    while (half >= 4) {            // as many radix-8 passes as fit; a radix-4 or radix-2 pass finishes the small spans
        fft_dif_pass8<SIGN>(a, n, half, tw);
        half >>= 3;
    }
#endif
    while (half >= 2) {
        const int q = half >> 1;
        const int lq = __ffs(q) - 1;                       // sizes are powers of two: shifts instead of divisions
        for (int t = threadIdx.x; t < (n >> 2); t += blockDim.x) {
            const int pos = t & (q - 1);
            const int i0 = ((t >> lq) << (lq + 2)) + pos;
            double2 x0 = a[SWZ(i0)], x1 = a[SWZ(i0 + q)], x2 = a[SWZ(i0 + half)], x3 = a[SWZ(i0 + half + q)];
            const double2 w1 = twiddle<SIGN>(tw, pos, 2 * half);
            const double2 w2 = cmul(w1, w1);                  // exp(i a)^2 instead of a second table look-up
            double2 s0 = cadd(x0, x2), d0 = cmul(csub(x0, x2), w1);
            double2 s1 = cadd(x1, x3), d1 = cmul(quarter_turn<SIGN>(csub(x1, x3)), w1);
            a[SWZ(i0)] = cadd(s0, s1);
            a[SWZ(i0 + q)] = cmul(csub(s0, s1), w2);
            a[SWZ(i0 + half)] = cadd(d0, d1);
            a[SWZ(i0 + half + q)] = cmul(csub(d0, d1), w2);
        }
        __syncthreads();
        half >>= 2;
    }
    if (half == 1) {
        for (int b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
            double2 u = a[SWZ(2 * b)], v = a[SWZ(2 * b + 1)];
            a[SWZ(2 * b)] = cadd(u, v);
            a[SWZ(2 * b + 1)] = csub(u, v);
        }
        __syncthreads();
    }
}
template <int SIGN>
__device__ __forceinline__ void fft_dit(double2* a, int n, const double2* __restrict__ tw) {
    int logn = 0;
    while ((1 << logn) < n) logn++;
    int half = 1;
#if FFT_RADIX8
    {   // mirror of fft_dif: the small-span remainder (radix-2 or radix-4) first, radix-8 passes for everything above
        const int rem = logn % 3;
        if (rem == 1) {
            for (int b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
                double2 u = a[SWZ(2 * b)], v = a[SWZ(2 * b + 1)];
                a[SWZ(2 * b)] = cadd(u, v);
                a[SWZ(2 * b + 1)] = csub(u, v);
            }
            __syncthreads();
            half = 2;
        } else if (rem == 2) {
            for (int t = threadIdx.x; t < (n >> 2); t += blockDim.x) {
                const int i0 = t << 2;
                double2 x0 = a[SWZ(i0)], x1 = a[SWZ(i0 + 1)], x2 = a[SWZ(i0 + 2)], x3 = a[SWZ(i0 + 3)];
                double2 y0 = cadd(x0, x1), y1 = csub(x0, x1), y2 = cadd(x2, x3), y3 = quarter_turn<SIGN>(csub(x2, x3));
                a[SWZ(i0)] = cadd(y0, y2);
                a[SWZ(i0 + 2)] = csub(y0, y2);
                a[SWZ(i0 + 1)] = cadd(y1, y3);
                a[SWZ(i0 + 3)] = csub(y1, y3);
            }
            __syncthreads();
            half = 4;
        }
        while (half < n) {
            fft_dit_pass8<SIGN>(a, n, half, tw);
            half <<= 3;
        }
        return;
    }
#endif
    if (logn & 1) {
        for (int b = threadIdx.x; b < (n >> 1); b += blockDim.x) {
            double2 u = a[SWZ(2 * b)], v = a[SWZ(2 * b + 1)];
            a[SWZ(2 * b)] = cadd(u, v);
            a[SWZ(2 * b + 1)] = csub(u, v);
        }
        __syncthreads();
        half = 2;
    }
    while (half < n) {
        const int h = half;
        const int lh = __ffs(h) - 1;
        for (int t = threadIdx.x; t < (n >> 2); t += blockDim.x) {
            const int pos = t & (h - 1);
            const int i0 = ((t >> lh) << (lh + 2)) + pos;
            double2 x0 = a[SWZ(i0)], x1 = a[SWZ(i0 + h)], x2 = a[SWZ(i0 + 2 * h)], x3 = a[SWZ(i0 + 3 * h)];
            const double2 wB = twiddle<SIGN>(tw, pos, 4 * h);
            if (h > 1) {
                const double2 wA = cmul(wB, wB);
                x1 = cmul(x1, wA);
                x3 = cmul(x3, wA);
            }
            double2 y0 = cadd(x0, x1), y1 = csub(x0, x1), y2 = cadd(x2, x3), y3 = csub(x2, x3);
            double2 u2 = cmul(y2, wB), u3 = cmul(quarter_turn<SIGN>(y3), wB);
            a[SWZ(i0)] = cadd(y0, u2);
            a[SWZ(i0 + 2 * h)] = csub(y0, u2);
            a[SWZ(i0 + h)] = cadd(y1, u3);
            a[SWZ(i0 + 3 * h)] = csub(y1, u3);
        }
        __syncthreads();
        half <<= 2;
    }
}

// Packed real FFT helpers.  `a` holds M = N/2 complex points z[n] = x[2n] + i x[2n+1]; after fft_dif<-1> the
// spectrum Z sits bit-reversed.  For bin k in [0, M]:  X[k] = (Z[k] + conj Z[M-k])/2 - (i/2) w^k (Z[k] - conj Z[M-k]),
// w = exp(-2 pi i / N).  power_of_pair returns |X[k]|^2 and |X[M-k]|^2.
__device__ __forceinline__ void real_bins_from_packed(double2 zk, double2 zmk, double2 wk, double2* xk, double2* xmk) {
    // e = (Z[k] + conj Z[M-k])/2 ; o = (Z[k] - conj Z[M-k])/2
    double ex = 0.5 * (zk.x + zmk.x), ey = 0.5 * (zk.y - zmk.y);
    double ox = 0.5 * (zk.x - zmk.x), oy = 0.5 * (zk.y + zmk.y);
    // -i * w^k * o
    double2 wo = cmul(wk, make_double2(ox, oy));
    *xk = make_double2(ex + wo.y, ey - wo.x);
    // X[M-k] = conj(e) - (-i w^k o)^*  ... derived from Hermitian symmetry: X[M-k] = conj(e) + i conj(w^k o)... use direct form
    *xmk = make_double2(ex - wo.y, -ey - wo.x);
}

// After fft_dif<-1>(a, M): replace the packed spectrum by the packed half-spectrum Y of a real-even spectrum
// F(P[k]) so that fft_dit<+1>(a, M) yields y[n] = c[2n] + i c[2n+1] with c = sum_{k<N} F(P)[k] exp(+2 pi i k n / N)
// (unnormalised inverse).  F is applied to the power |X[k]|^2 (identity for autocorrelation, log for the cepstrum).
// If pw != nullptr the (transformed) half-spectrum values F(P[k]), k = 0..M, are also written there.
template <class F>
__device__ __forceinline__ void packed_power_to_inverse_input(double2* a, int M, int logM, const double2* __restrict__ tw,
                                                              F f, double* pw) {
    int N = 2 * M;
    for (int k = threadIdx.x; k <= M / 2; k += blockDim.x) {
        if (k == 0) {
            double2 z0 = a[SWZ(0)];
            double p0 = f((z0.x + z0.y) * (z0.x + z0.y));
            double pM = f((z0.x - z0.y) * (z0.x - z0.y));
            a[SWZ(0)] = make_double2(p0 + pM, p0 - pM);
            if (pw) { pw[0] = p0; pw[M] = pM; }
        } else {
            int ik = bitrev(k, logM), imk = bitrev(M - k, logM);
            double2 zk = a[SWZ(ik)], zmk = a[SWZ(imk)];
            double2 wk = __ldg(tw + k * (TW_N / N));           // exp(-2 pi i k / N)
            double2 xk, xmk;
            real_bins_from_packed(zk, zmk, wk, &xk, &xmk);
            double pk = f(xk.x * xk.x + xk.y * xk.y);
            double pmk = f(xmk.x * xmk.x + xmk.y * xmk.y);
            if (pw) { pw[k] = pk; pw[M - k] = pmk; }
            // Y[k] = (P[k] + P[M-k]) + i w'^k (P[k] - P[M-k]),  w' = conj(w);  Y[M-k] likewise with k -> M-k
            double s = pk + pmk, d = pk - pmk;
            // i * conj(wk) * d = i*(wk.x - i wk.y)*d = (wk.y*d) + i (wk.x*d)
            double2 yk = make_double2(s + wk.y * d, wk.x * d);
            // w'^(M-k) = exp(2 pi i (M-k)/N) = exp(i pi) * conj(w'^k)... = -(wk.x + i wk.y) ; P diff = -d
            // i * (-(wk.x + i wk.y)) * (-d) = i (wk.x + i wk.y) d = (-wk.y d) + i (wk.x d)
            double2 ymk = make_double2(s - wk.y * d, wk.x * d);
            a[SWZ(ik)] = yk;
            if (imk != ik) a[SWZ(imk)] = ymk;
        }
    }
    __syncthreads();
}

struct IdentityF { __device__ __forceinline__ double operator()(double p) const { return p; } };
