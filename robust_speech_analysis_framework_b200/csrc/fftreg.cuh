// fftreg.cuh -- register-resident float64 FFT building blocks for the warp-per-frame transforms (k_acw.cu, ...).
//
// A frame's packed real transform (M = 512 or 1024 complex points) is split Cooley-Tukey style as M = L x 32 with
// L = M / 32 lanes per frame (a whole warp for M = 1024, half a warp for M = 512 -- two frames per warp):
//
//   pass A   lane j holds z[j + L k], k = 0..31, in registers: a 32-point FFT without leaving the register file
//   twiddle  times W_M^(j q)                      (one coalesced table row per q)
//   exchange ONE trip through shared memory (XOR-swizzled [q][j] layout, conflict-free both ways, __syncwarp only)
//   pass B   32 rows of L-point FFTs, one (L = 32) or two (L = 16) rows per lane, again in registers
//
// so that lane j ends up holding Z[j + L r], r = 0..31 -- the same "stride L" layout pass A starts from.  The real-input
// untangle / power / retangle step pairs bin k with M - k, which live in lanes j and L - j: it is done with warp shuffles
// (each lane computes half of its pairs and trades the results), and the inverse transform starts from the registers the
// forward one ended in.  Per transform the data crosses shared memory once instead of four times (radix-8 passes of
// fft.cuh) and no block-wide barrier is executed at all.
//
// Everything here is plain C++ on statically indexed arrays (template recursion => every index is a compile-time
// constant => arrays stay in registers) and also compiles for the host, where tests/fftreg_host_test.cpp checks it
// against a direct DFT.  Arithmetic of the reference these transforms serve: Praat's NUMfft (fon/Sound_to_Pitch.cpp
// autocorrelation, mshds_extractor.py:104,143,178,270,355), same DFT, different operation order.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#else
#define FR_HD inline
#ifndef FR_HAVE_DOUBLE2
#define FR_HAVE_DOUBLE2
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#endif
#endif

// cos / sin of 2 pi k / 64, k = 0..16 (a quarter circle); the other values follow by symmetry at compile time
FR_HD constexpr double fr_cosq(int k) {
    return k == 0 ? 1.0 : k == 1 ? 0.99518472667219688624 : k == 2 ? 0.98078528040323044913 : k == 3 ? 0.95694033573220886494 :
           k == 4 ? 0.92387953251128675613 : k == 5 ? 0.88192126434835502971 : k == 6 ? 0.83146961230254523708 :
           k == 7 ? 0.77301045336273696081 : k == 8 ? 0.70710678118654752440 : k == 9 ? 0.63439328416364549822 :
           k == 10 ? 0.55557023301960222474 : k == 11 ? 0.47139673682599764856 : k == 12 ? 0.38268343236508977173 :
           k == 13 ? 0.29028467725446236764 : k == 14 ? 0.19509032201612826785 : k == 15 ? 0.09801714032956060199 : 0.0;
}
// cos(2 pi k / 64) for any k
FR_HD constexpr double fr_cos64(int k) {
    k = ((k % 64) + 64) % 64;
    return k <= 16 ? fr_cosq(k) : k <= 32 ? -fr_cosq(32 - k) : k <= 48 ? -fr_cosq(k - 32) : fr_cosq(64 - k);
}
FR_HD constexpr double fr_sin64(int k) { return fr_cos64(k - 16); }

FR_HD constexpr int fr_brev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++) r = (r << 1) | ((v >> i) & 1);
    return r;
}

FR_HD double2 fr_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
FR_HD double2 fr_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
FR_HD double2 fr_mul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)); }

// a * exp(SIGN * 2 pi i K64 / 64) with the trivial angles folded at compile time
template <int K64, int SIGN>
FR_HD double2 fr_twmul(double2 a) {
    constexpr int K = ((K64 % 64) + 64) % 64;
    if constexpr (K == 0) return a;
    else if constexpr (K == 16) return SIGN > 0 ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
    else if constexpr (K == 32) return make_double2(-a.x, -a.y);
    else if constexpr (K == 48) return SIGN > 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
    else if constexpr (K == 8) {
        constexpr double r = 0.70710678118654752440;
        return SIGN > 0 ? make_double2((a.x - a.y) * r, (a.x + a.y) * r) : make_double2((a.x + a.y) * r, (a.y - a.x) * r);
    } else if constexpr (K == 24) {
        constexpr double r = 0.70710678118654752440;
        return SIGN > 0 ? make_double2(-(a.x + a.y) * r, (a.x - a.y) * r) : make_double2((a.y - a.x) * r, -(a.x + a.y) * r);
    } else {
        constexpr double c = fr_cos64(K), s = (SIGN > 0 ? 1.0 : -1.0) * fr_sin64(K);
        return make_double2(fma(a.x, c, -(a.y * s)), fma(a.x, s, a.y * c));
    }
}

// ---- radix-2 decimation in frequency on N registers: natural order in, bit-reversed order out (X[k] at x[brev(k)]) ----
template <int N, int HALF, int SIGN, int I>
struct FrDifStage {
    static FR_HD void run(double2* x) {
        constexpr int p = I % HALF, g = I / HALF, i0 = g * 2 * HALF + p;
        const double2 u = x[i0], v = x[i0 + HALF];
        x[i0] = fr_add(u, v);
        x[i0 + HALF] = fr_twmul<p * (64 / (2 * HALF)), SIGN>(fr_sub(u, v));
        FrDifStage<N, HALF, SIGN, I + 1>::run(x);
    }
};
template <int N, int HALF, int SIGN>
struct FrDifStage<N, HALF, SIGN, N / 2> {
    static FR_HD void run(double2*) {}
};
template <int N, int HALF, int SIGN>
FR_HD void fr_dif_stages(double2* x) {
    FrDifStage<N, HALF, SIGN, 0>::run(x);
    if constexpr (HALF > 1) fr_dif_stages<N, HALF / 2, SIGN>(x);
}
// N in {8, 16, 32}; SIGN = -1 forward, +1 inverse (unnormalised)
template <int N, int SIGN>
FR_HD void fr_fft(double2* x) { fr_dif_stages<N, N / 2, SIGN>(x); }

// compile-time loop: f(FrInt<I>) for I = BEGIN .. END-1, so that every array index derived from the loop variable is a
// constant expression (register arrays must never see a run-time index)
template <int V> struct FrInt { static constexpr int value = V; };
template <int BEGIN, int END, class F>
FR_HD void fr_static_for(F&& f) {
    if constexpr (BEGIN < END) {
        f(FrInt<BEGIN>{});
        fr_static_for<BEGIN + 1, END>(f);
    }
}

// register slot of logical element r (k = j + L r) after pass B
template <int L>
FR_HD constexpr int fr_slot(int r) { return L == 32 ? fr_brev(r, 5) : 16 * (r & 1) + fr_brev(r >> 1, 4); }

// shared-memory position (in complex elements, frame-local) of pass-A output (q, j): XOR swizzle of the lane index
template <int L>
FR_HD constexpr int fr_xch(int q, int j) { return q * L + (j ^ (q & 7)); }

// Real-input untangle split in two: the powers |X[k]|^2, |X[M-k]|^2 of a bin pair of the packed transform ...
FR_HD void fr_pair_powers(double2 zk, double2 zmk, double2 wk, double* pk_out, double* pmk_out) {
    const double ex = 0.5 * (zk.x + zmk.x), ey = 0.5 * (zk.y - zmk.y);
    const double ox = 0.5 * (zk.x - zmk.x), oy = 0.5 * (zk.y + zmk.y);
    const double2 wo = fr_mul(wk, make_double2(ox, oy));
    const double xkx = ex + wo.y, xky = ey - wo.x;
    const double xmx = ex - wo.y, xmy = -ey - wo.x;
    *pk_out = fma(xkx, xkx, xky * xky);
    *pmk_out = fma(xmx, xmx, xmy * xmy);
}
// ... and the packed inverse-transform input of a real, even spectrum G: Y[k] = (G[k] + G[M-k]) + i conj(w^k) (G[k] - G[M-k])
FR_HD void fr_pair_retangle(double gk, double gmk, double2 wk, double2* yk, double2* ymk) {
    const double s = gk + gmk, d = gk - gmk;
    *yk = make_double2(fma(wk.y, d, s), wk.x * d);
    *ymk = make_double2(fma(-wk.y, d, s), wk.x * d);
}

// Real-input untangle for one bin pair.  Zk = Z[k], Zmk = Z[M - k] of the packed transform, wk = exp(-2 pi i k / N), N = 2M.
// Returns through yk / ymk the packed input of the inverse transform of the (real, even) power spectrum:
// Y[k] = (P[k] + P[M-k]) + i conj(w^k) (P[k] - P[M-k]) and Y[M-k]; pk / pmk receive the powers |X[k]|^2, |X[M-k]|^2.
FR_HD void fr_pair(double2 zk, double2 zmk, double2 wk, double2* yk, double2* ymk, double* pk_out, double* pmk_out) {
    const double ex = 0.5 * (zk.x + zmk.x), ey = 0.5 * (zk.y - zmk.y);
    const double ox = 0.5 * (zk.x - zmk.x), oy = 0.5 * (zk.y + zmk.y);
    const double2 wo = fr_mul(wk, make_double2(ox, oy));
    const double xkx = ex + wo.y, xky = ey - wo.x;
    const double xmx = ex - wo.y, xmy = -ey - wo.x;
    const double pk = fma(xkx, xkx, xky * xky), pmk = fma(xmx, xmx, xmy * xmy);
    const double s = pk + pmk, d = pk - pmk;
    *yk = make_double2(fma(wk.y, d, s), wk.x * d);
    *ymk = make_double2(fma(-wk.y, d, s), wk.x * d);
    *pk_out = pk; *pmk_out = pmk;
}
