// num.cuh -- warp-cooperative restatement of Praat's sinc interpolation and Brent peak refinement
// (melder/NUMinterpol.cpp NUM_interpolate_sinc / NUMimproveExtremum, dwsys/NUM2.cpp NUMminimize_brent), the
// arithmetic behind every to_pitch_* / to_harmonicity_cc call in mshds_extractor.py and the Sinc70 extrema at :78.
//
// One warp evaluates one interpolation: the 2*depth taps are split over the 32 lanes and combined with a fixed-order
// butterfly sum, so results are deterministic.  The Brent state machine is replicated in every lane (no divergence).
#pragma once
#include "common.cuh"

// 1-based accessors: a plain pointer, or a symmetric correlation row stored for lags 0..len-1 only
struct PtrY {
    const double* y;
    __device__ __forceinline__ double operator()(int j) const { return y[j]; }
    __device__ __forceinline__ int lo() const { return 1; }              // entries outside [lo, hi] are known to be zero
    __device__ __forceinline__ int hi(int n) const { return n; }
};
struct SymRowY {            // y(j) = r[|j - centre|], zero beyond the stored lags
    const double* __restrict__ row;
    int centre, len;
    __device__ __forceinline__ double operator()(int j) const {
        int l = j - centre;
        l = l < 0 ? -l : l;
        return l < len ? row[l] : 0.0;
    }
    __device__ __forceinline__ int lo() const { return centre - len + 1; }
    __device__ __forceinline__ int hi(int) const { return centre + len - 1; }
};

// cos(pi*u) for u in [0, 1] from the FFT twiddle table (tw[j] = (cos, -sin)(pi*j/4096), j < 4096) plus a 4th-order
// angle-addition correction: |error| ~ 1e-16, about half the float64 instructions of cospi().
__device__ __forceinline__ double cospi_tab(double u, const double2* __restrict__ tw) {
    const double magic = 6755399441055744.0;                 // 2^52 + 2^51: round-to-nearest-integer trick
    double t = fma(u, 4096.0, magic);
    int j = __double2loint(t);
    double jd = t - magic;
    if (j > 4095) { j = 4095; jd = 4095.0; }
    double x = (u - jd * (1.0 / 4096.0)) * MSHDS_PI;         // |x| <= pi/8192 (a little more at u = 1)
    double x2 = x * x;
    double2 T = __ldg(tw + j);
    double cx = fma(x2, fma(x2, 1.0 / 24.0, -0.5), 1.0);
    double sx = x * fma(x2, -1.0 / 6.0, 1.0);
    return fma(T.x, cx, T.y * sx);
}

struct StagedY {            // a window of the 1-based array staged in shared memory: y(j) = st[j - j0]
    const double* st;
    int j0;
    __device__ __forceinline__ double operator()(int j) const { return st[j - j0]; }
    __device__ __forceinline__ int lo() const { return 1; }
    __device__ __forceinline__ int hi(int n) const { return n; }
};

// y is 1-based: y(1..n) valid.  All 32 lanes must call with identical arguments; all lanes get the result.
// A group of `nl` (32 or 16) adjacent lanes, all named in `mask`, cooperates; groups of one warp may diverge.
__device__ __forceinline__ double group_sum(double v, unsigned mask, int nl) {
    for (int o = nl >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

template <class Y>
__device__ __forceinline__ double sinc_interp_warp_t(const Y& y, int n, double x, int maxDepth, int lane,
                                                     const double2* __restrict__ tw, unsigned mask = FULL_MASK, int nl = 32) {
    int midleft = (int)floor(x), midright = midleft + 1;
    if (n < 1) return DEVNAN;
    if (x > n) return y(n);
    if (x < 1) return y(1);
    if (x == midleft) return y(midleft);
    if (maxDepth > midright - 1) maxDepth = midright - 1;
    if (maxDepth > n - midleft) maxDepth = n - midleft;
    if (maxDepth <= 0) return y((int)floor(x + 0.5));
    if (maxDepth == 1) return y(midleft) + (x - midleft) * (y(midright) - y(midleft));
    if (maxDepth == 2) {
        double yl = y(midleft), yr = y(midright);
        double dyl = 0.5 * (yr - y(midleft - 1)), dyr = 0.5 * (y(midright + 1) - yl);
        double fil = x - midleft, fir = midright - x;
        return yl * fir + yr * fil - fil * fir * (0.5 * (dyr - dyl) + (fil - 0.5) * (dyl + dyr - 2 * (yr - yl)));
    }
    // left side: taps ix = midleft - k, k = 0..maxDepth-1; a_k = pi*(x-midleft) + k*pi; aa_k = a_k / (x-left+1)
    // right side: ix = midright + k;  a_k = pi*(midright-x) + k*pi; aa_k = a_k / (right-x+1)
    double fl = x - midleft, fr = midright - x;
    // halfsina / a * (1 + cos(aa)) with a = pi*(f + k): the constant factor 0.5*sin(pi f)/pi is hoisted and the
    // per-tap division becomes an IEEE reciprocal (one rounding more than Praat's expression: ~1 ulp per tap)
    // sin(pi*fr) = sin(pi*(1 - fl)) = sin(pi*fl) = cos(pi*|fl - 1/2|): one table look-up serves both sides
    const double hsl = 0.5 * cospi_tab(fabs(fl - 0.5), tw) * (1.0 / MSHDS_PI), hsr = hsl;
    double invl = __drcp_rn(fl + maxDepth), invr = __drcp_rn(fr + maxDepth);  // x-left+1 = fl + depth ; right-x+1 = fr + depth
    double accl = 0.0, accr = 0.0;
    // window 1 + cos(pi*(f + k)/(f + depth)) for this lane's taps k = lane, lane + nl, ...: the angles advance by a constant
    // step, so after two table look-ups per side the cosines follow the Chebyshev recurrence c[m+1] = 2 cos(step) c[m] - c[m-1]
    const double stepl = (double)nl * invl, stepr = (double)nl * invr;
    const bool rec = stepl <= 1.0 && stepr <= 1.0;
    const double twocl = rec ? 2.0 * cospi_tab(stepl, tw) : 0.0, twocr = rec ? 2.0 * cospi_tab(stepr, tw) : 0.0;
    double clm = 0.0, cl = 0.0, crm = 0.0, cr = 0.0;
    // taps that only meet entries known to be zero (a correlation row is stored up to its last useful lag; the 700-deep
    // window of a maximum at lag 2-3 reaches twice as far) add exactly nothing: stop at the last tap that can matter
    {
        const int kl = midleft - y.lo() + 1, kr = y.hi(n) - midright + 1;
        const int kmax = kl > kr ? kl : kr;
        if (kmax < maxDepth) maxDepth = kmax > 0 ? kmax : 0;
    }
    // two taps per side and trip: the four divisions 1/(f + k) share ONE reciprocal (of the product of the four
    // denominators, <= 71^4), the individual ones are recovered with multiplications
    int k = lane, m = 0;
    for (; k + nl < maxDepth; k += 2 * nl, m += 2) {
        const int k2 = k + nl;
        const double al = fl + k, ar = fr + k;                              // in units of pi
        const double al2 = fl + k2, ar2 = fr + k2;
        double cl0, cr0, cl1, cr1;
        if (!rec || m < 2) {
            cl0 = cospi_tab(fmin(al * invl, 1.0), tw);
            cr0 = cospi_tab(fmin(ar * invr, 1.0), tw);
            cl1 = cospi_tab(fmin(al2 * invl, 1.0), tw);
            cr1 = cospi_tab(fmin(ar2 * invr, 1.0), tw);
        } else {
            cl0 = fma(twocl, cl, -clm);
            cr0 = fma(twocr, cr, -crm);
            cl1 = fma(twocl, cl0, -cl);
            cr1 = fma(twocr, cr0, -cr);
        }
        clm = cl0; cl = cl1; crm = cr0; cr = cr1;
        const double p = al * ar, q = al2 * ar2;
        const double r = __drcp_rn(p * q);
        const double ip = q * r, iq = p * r;
        const double dl = (ar * ip) * (1.0 + cl0), dr = (al * ip) * (1.0 + cr0);
        const double dl2 = (ar2 * iq) * (1.0 + cl1), dr2 = (al2 * iq) * (1.0 + cr1);
        accl = fma(y(midleft - k), dl, accl);
        accr = fma(y(midright + k), dr, accr);
        accl = fma(y(midleft - k2), dl2, accl);
        accr = fma(y(midright + k2), dr2, accr);
    }
    if (k < maxDepth) {                                                     // odd tap left over
        const double al = fl + k, ar = fr + k;
        double cl0, cr0;
        if (!rec || m < 2) {
            cl0 = cospi_tab(fmin(al * invl, 1.0), tw);
            cr0 = cospi_tab(fmin(ar * invr, 1.0), tw);
        } else {
            cl0 = fma(twocl, cl, -clm);
            cr0 = fma(twocr, cr, -crm);
        }
        const double r = __drcp_rn(al * ar);
        accl = fma(y(midleft - k), (ar * r) * (1.0 + cl0), accl);
        accr = fma(y(midright + k), (al * r) * (1.0 + cr0), accr);
    }
    // (-1)^k of the tap: k = lane + j * nl has the parity of the lane (nl is even), so the sign is applied once to the
    // lane's sums instead of to every sample
    if (lane & 1) { accl = -accl; accr = -accr; }
    return group_sum(accl * hsl + accr * hsr, mask, nl);
}

__device__ __forceinline__ double sinc_interp_warp(const double* y, int n, double x, int maxDepth, int lane,
                                                   const double2* __restrict__ tw) {
    PtrY a{y};
    return sinc_interp_warp_t(a, n, x, maxDepth, lane, tw);
}

// NUMminimize_brent specialised to f(x) = -/+ sinc_interp(y, x, depth); returns x of the extremum, *fx its value
// (already sign-corrected: the interpolated y at the extremum).
template <class Y>
__device__ __forceinline__ double brent_sinc_warp_t(const Y& y, int n, double a, double b, int depth, bool isMaximum,
                                                    double* fx_out, int lane, const double2* __restrict__ tw,
                                                    unsigned mask = FULL_MASK, int nl = 32) {
    const double golden = 1.0 - 0.6180339887498948482045868343656381177203;
    const double sqrt_epsilon = 1.4901161193847656e-08;   // sqrt(DBL_EPSILON)
    const double tol = 1e-10;
    const double sg = isMaximum ? -1.0 : 1.0;
    double x, v, fv, w, fw, fx;
    v = a + golden * (b - a);
    fv = sg * sinc_interp_warp_t(y, n, v, depth, lane, tw, mask, nl);
    x = v; w = v;
    fx = fv; fw = fv;
    for (int iter = 1; iter <= 60; iter++) {
        double range = b - a;
        double middle_range = (a + b) / 2.0;
        double tol_act = sqrt_epsilon * fabs(x) + tol / 3.0;
        double new_step;
        if (fabs(x - middle_range) + range / 2.0 <= 2.0 * tol_act) break;
        new_step = golden * (x < middle_range ? b - x : a - x);
        if (fabs(x - w) >= tol_act) {
            double p, q, t;
            t = (x - w) * (fx - fv);
            q = (x - v) * (fx - fw);
            p = (x - v) * q - (x - w) * t;
            q = 2.0 * (q - t);
            if (q > 0.0) p = -p; else q = -q;
            if (fabs(p) < fabs(new_step * q) && p > q * (a - x + 2.0 * tol_act) && p < q * (b - x - 2.0 * tol_act))
                new_step = p / q;
        }
        if (fabs(new_step) < tol_act) new_step = new_step > 0.0 ? tol_act : -tol_act;
        {
            double t = x + new_step;
            double ft = sg * sinc_interp_warp_t(y, n, t, depth, lane, tw, mask, nl);
            if (ft <= fx) {
                if (t < x) b = x; else a = x;
                v = w; w = x; x = t;
                fv = fw; fw = fx; fx = ft;
            } else {
                if (t < x) a = t; else b = t;
                if (ft <= fw || w == x) {
                    v = w; w = t;
                    fv = fw; fw = ft;
                } else if (ft <= fv || v == x || v == w) {
                    v = t;
                    fv = ft;
                }
            }
        }
    }
    *fx_out = sg * fx;
    return x;
}

#define PEAK_NONE 0
#define PEAK_PARABOLIC 1
#define PEAK_SINC70 3
#define PEAK_SINC700 4

// NUMimproveExtremum (warp-cooperative for the sinc modes). y 1-based.
template <class Y>
__device__ __forceinline__ double improve_extremum_warp_t(const Y& y, int n, int ixmid, int interpolation,
                                                          double* ixmid_real, bool isMaximum, int lane,
                                                          const double2* __restrict__ tw, unsigned mask = FULL_MASK,
                                                          int nl = 32) {
    if (ixmid <= 1) { *ixmid_real = 1; return y(1); }
    if (ixmid >= n) { *ixmid_real = n; return y(n); }
    if (interpolation <= PEAK_NONE) { *ixmid_real = ixmid; return y(ixmid); }
    if (interpolation == PEAK_PARABOLIC) {
        double dy = 0.5 * (y(ixmid + 1) - y(ixmid - 1));
        double d2y = 2 * y(ixmid) - y(ixmid - 1) - y(ixmid + 1);
        *ixmid_real = ixmid + dy / d2y;
        return y(ixmid) + 0.5 * dy * dy / d2y;
    }
    double fx;
    *ixmid_real = brent_sinc_warp_t(y, n, (double)(ixmid - 1), (double)(ixmid + 1), interpolation == PEAK_SINC70 ? 70 : 700,
                                    isMaximum, &fx, lane, tw, mask, nl);
    return fx;
}
__device__ __forceinline__ double improve_extremum_warp(const double* y, int n, int ixmid, int interpolation,
                                                        double* ixmid_real, bool isMaximum, int lane,
                                                        const double2* __restrict__ tw) {
    PtrY a{y};
    return improve_extremum_warp_t(a, n, ixmid, interpolation, ixmid_real, isMaximum, lane, tw);
}

// single-thread parabolic / none variant
__device__ __forceinline__ double improve_extremum_simple(const double* y, int n, int ixmid, int interpolation,
                                                          double* ixmid_real) {
    if (ixmid <= 1) { *ixmid_real = 1; return y[1]; }
    if (ixmid >= n) { *ixmid_real = n; return y[n]; }
    if (interpolation <= PEAK_NONE) { *ixmid_real = ixmid; return y[ixmid]; }
    double dy = 0.5 * (y[ixmid + 1] - y[ixmid - 1]);
    double d2y = 2 * y[ixmid] - y[ixmid - 1] - y[ixmid + 1];
    *ixmid_real = ixmid + dy / d2y;
    return y[ixmid] + 0.5 * dy * dy / d2y;
}

// NUMbessel_i0_f (Abramowitz & Stegun 9.8.1 / 9.8.2), used by the Kaiser-20 intensity window
__host__ __device__ inline double bessel_i0_f(double x) {
    if (x < 0.0) x = -x;
    if (x < 3.75) {
        double t = x / 3.75;
        t *= t;
        return 1.0 + t * (3.5156229 + t * (3.0899424 + t * (1.2067492 + t * (0.2659732 + t * (0.0360768 + t * 0.0045813)))));
    }
    double t = 3.75 / x;
    return exp(x) / sqrt(x) *
           (0.39894228 + t * (0.01328592 + t * (0.00225319 + t * (-0.00157565 + t * (0.00916281 +
            t * (-0.02057706 + t * (0.02635537 + t * (-0.01647633 + t * 0.00392377))))))));
}
