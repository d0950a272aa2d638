/*
 * oracle/praat_num.c -- TEST INFRASTRUCTURE (see praat_core.h header).  PARITY UNPINNED.
 *
 * Restates Praat 6.1.38 numerics used by every analysis behind /root/reference/src/mshds_extractor.py:
 *   melder/NUMinterpol (NUM_interpolate_sinc, NUMimproveExtremum), dwsys/NUM2 (NUMminimize_brent,
 *   NUMlineFit_theil), melder/NUM (NUMquantile), num/NUMspecfunc (NUMbessel_i0_f), fon/Sampled.cpp
 *   (Sampled_shortTermAnalysis, Sampled_getWindowSamples), fon/Vector.cpp (value/min/max queries).
 */
#include "praat_core.h"
#include <float.h>

/* ---------------- Sound / Sampled ---------------- */

Sound *sound_create(double xmin, double xmax, long nx, double dx, double x1) {
    Sound *s = (Sound *)calloc(1, sizeof(Sound));
    s->xmin = xmin; s->xmax = xmax; s->nx = nx; s->dx = dx; s->x1 = x1;
    s->z = (double *)calloc((size_t)(nx > 0 ? nx : 1), sizeof(double));
    s->owns = 1;
    return s;
}

/* parselmouth.Sound(filepath) for a 16-bit mono WAV (mshds_extractor.py:415): samples / 32768,
 * xmin = 0, xmax = nx*dx, x1 = dx/2 (fon/Sound.cpp Sound_createSimple). */
Sound *sound_from_pcm16(const int16_t *pcm, long n, double fs) {
    double dx = 1.0 / fs;
    Sound *s = sound_create(0.0, n * dx, n, dx, 0.5 * dx);
    for (long i = 0; i < n; i++) s->z[i] = (double)pcm[i] / 32768.0;
    return s;
}

Sound *sound_copy(const Sound *me) {
    Sound *s = sound_create(me->xmin, me->xmax, me->nx, me->dx, me->x1);
    memcpy(s->z, me->z, (size_t)me->nx * sizeof(double));
    return s;
}

void sound_free(Sound *s) {
    if (!s) return;
    if (s->owns) free(s->z);
    free(s);
}

void contour_free(Contour *c) {
    if (!c) return;
    free(c->y);
    free(c);
}

/* fon/Sampled.cpp Sampled_shortTermAnalysis */
int shortTermAnalysis(long nx, double dx, double x1, double windowDuration, double timeStep, long *numberOfFrames,
                      double *firstTime) {
    volatile double myDuration = dx * nx;
    if (windowDuration > myDuration) return 0;
    *numberOfFrames = (long)floor((myDuration - windowDuration) / timeStep) + 1;
    double ourMidTime = x1 - 0.5 * dx + 0.5 * myDuration;
    double thyDuration = *numberOfFrames * timeStep;
    *firstTime = ourMidTime - 0.5 * thyDuration + 0.5 * timeStep;
    return 1;
}

/* fon/Sampled.cpp Sampled_getWindowSamples */
long getWindowSamples(double x1, double dx, long nx, double xmin, double xmax, long *ixmin, long *ixmax) {
    double rixmin = 1.0 + ceil((xmin - x1) / dx);
    double rixmax = 1.0 + floor((xmax - x1) / dx);
    *ixmin = rixmin < 1.0 ? 1 : (long)rixmin;
    *ixmax = rixmax > (double)nx ? nx : (long)rixmax;
    if (*ixmin > *ixmax) return 0;
    return *ixmax - *ixmin + 1;
}

/* ---------------- special functions ---------------- */

/* num/NUMspecfunc NUMbessel_i0_f: Abramowitz & Stegun 9.8.1 / 9.8.2 polynomial approximations. */
double NUMbessel_i0_f(double x) {
    if (x < 0.0) return NUMbessel_i0_f(-x);
    if (x < 3.75) {
        double t = x / 3.75;
        t *= t;
        return 1.0 + t * (3.5156229 + t * (3.0899424 + t * (1.2067492 + t * (0.2659732 + t * (0.0360768 + t * 0.0045813)))));
    }
    double t = 3.75 / x;
    return exp(x) / sqrt(x) *
           (0.39894228 +
            t * (0.01328592 +
                 t * (0.00225319 +
                      t * (-0.00157565 +
                           t * (0.00916281 + t * (-0.02057706 + t * (0.02635537 + t * (-0.01647633 + t * 0.00392377))))))));
}

/* ---------------- sinc interpolation ---------------- */

/* melder/NUMinterpol.cpp NUM_interpolate_sinc; y1 is 1-based [1..n]. */
double NUM_interpolate_sinc(const double *y, long n, double x, long maxDepth) {
    long midleft = (long)floor(x), midright = midleft + 1;
    double result = 0.0;
    if (n < 1) return UNDEF;
    if (x > n) return y[n];
    if (x < 1) return y[1];
    if (x == midleft) return y[midleft];
    if (maxDepth > midright - 1) maxDepth = midright - 1;
    if (maxDepth > n - midleft) maxDepth = n - midleft;
    if (maxDepth <= 0) return y[iround(x)];
    if (maxDepth == 1) return y[midleft] + (x - midleft) * (y[midright] - y[midleft]);
    if (maxDepth == 2) {
        double yl = y[midleft], yr = y[midright];
        double dyl = 0.5 * (yr - y[midleft - 1]), dyr = 0.5 * (y[midright + 1] - yl);
        double fil = x - midleft, fir = midright - x;
        return yl * fir + yr * fil - fil * fir * (0.5 * (dyr - dyl) + (fil - 0.5) * (dyl + dyr - 2 * (yr - yl)));
    }
    long left = midright - maxDepth, right = midleft + maxDepth;
    double a = NUMpi * (x - midleft);
    double halfsina = 0.5 * sin(a);
    double aa = a / (x - left + 1.0);
    double daa = NUMpi / (x - left + 1.0);
    for (long ix = midleft; ix >= left; ix--) {
        double d = halfsina / a * (1.0 + cos(aa));
        result += y[ix] * d;
        a += NUMpi;
        aa += daa;
        halfsina = -halfsina;
    }
    a = NUMpi * (midright - x);
    halfsina = 0.5 * sin(a);
    aa = a / (right - x + 1.0);
    daa = NUMpi / (right - x + 1.0);
    for (long ix = midright; ix <= right; ix++) {
        double d = halfsina / a * (1.0 + cos(aa));
        result += y[ix] * d;
        a += NUMpi;
        aa += daa;
        halfsina = -halfsina;
    }
    return result;
}

/* dwsys/NUM2.cpp NUMminimize_brent (netlib fminbr). */
double NUMminimize_brent(double (*f)(double, void *), double a, double b, void *closure, double tol, double *fx) {
    double x, v, fv, w, fw;
    const double golden = 1.0 - 0.6180339887498948482045868343656381177203;
    const double sqrt_epsilon = sqrt(DBL_EPSILON);
    const long itermax = 60;
    v = a + golden * (b - a);
    fv = f(v, closure);
    x = v; w = v;
    *fx = fv; fw = fv;
    for (long iter = 1; iter <= itermax; iter++) {
        double range = b - a;
        double middle_range = (a + b) / 2.0;
        double tol_act = sqrt_epsilon * fabs(x) + tol / 3.0;
        double new_step;
        if (fabs(x - middle_range) + range / 2.0 <= 2.0 * tol_act) return x;
        new_step = golden * (x < middle_range ? b - x : a - x);
        if (fabs(x - w) >= tol_act) {
            double p, q, t;
            t = (x - w) * (*fx - fv);
            q = (x - v) * (*fx - fw);
            p = (x - v) * q - (x - w) * t;
            q = 2.0 * (q - t);
            if (q > 0.0) p = -p; else q = -q;
            if (fabs(p) < fabs(new_step * q) && p > q * (a - x + 2.0 * tol_act) && p < q * (b - x - 2.0 * tol_act))
                new_step = p / q;
        }
        if (fabs(new_step) < tol_act) new_step = new_step > 0.0 ? tol_act : -tol_act;
        {
            double t = x + new_step;
            double ft = f(t, closure);
            if (ft <= *fx) {
                if (t < x) b = x; else a = x;
                v = w; w = x; x = t;
                fv = fw; fw = *fx; *fx = ft;
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
    return x;
}

struct improve_params { long depth; const double *y; long n; int isMaximum; };
static double improve_evaluate(double x, void *closure) {
    struct improve_params *me = (struct improve_params *)closure;
    double y = NUM_interpolate_sinc(me->y, me->n, x, me->depth);
    return me->isMaximum ? -y : y;
}

/* melder/NUMinterpol.cpp NUMimproveExtremum */
double NUMimproveExtremum(const double *y, long n, long ixmid, int interpolation, double *ixmid_real, int isMaximum) {
    struct improve_params params;
    double result;
    if (ixmid <= 1) { *ixmid_real = 1; return y[1]; }
    if (ixmid >= n) { *ixmid_real = n; return y[n]; }
    if (interpolation <= PEAK_NONE) { *ixmid_real = ixmid; return y[ixmid]; }
    if (interpolation == PEAK_PARABOLIC) {
        double dy = 0.5 * (y[ixmid + 1] - y[ixmid - 1]);
        double d2y = 2 * y[ixmid] - y[ixmid - 1] - y[ixmid + 1];
        *ixmid_real = ixmid + dy / d2y;
        return y[ixmid] + 0.5 * dy * dy / d2y;
    }
    params.y = y; params.n = n;
    params.depth = interpolation == PEAK_SINC70 ? 70 : 700;
    params.isMaximum = isMaximum;
    *ixmid_real = NUMminimize_brent(improve_evaluate, ixmid - 1, ixmid + 1, &params, 1e-10, &result);
    return isMaximum ? -result : result;
}

/* ---------------- sorting / quantiles / Theil ---------------- */

static int cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : x > y ? 1 : 0;
}
void sort_doubles(double *a, long n) { qsort(a, (size_t)n, sizeof(double), cmp_double); }

/* melder/NUM.cpp NUMquantile; a is sorted and 1-based */
double NUMquantile(const double *a, long n, double factor) {
    double place = factor * n + 0.5;
    long left = (long)floor(place);
    if (n < 1) return 0.0;
    if (left < 1) return a[1];
    if (left >= n) return a[n];
    if (a[left + 1] == a[left]) return a[left];
    return a[left] + (place - left) * (a[left + 1] - a[left]);
}

/* dwsys/NUM2.cpp NUMlineFit_theil; x,y 1-based.  "Robust" (method 2) = incomplete. */
void NUMlineFit_theil(const double *x, const double *y, long n, double *out_m, double *out_intercept, int complete) {
    double m, intercept;
    if (n <= 0) {
        m = intercept = UNDEF;
    } else if (n == 1) {
        intercept = y[1];
        m = 0.0;
    } else if (n == 2) {
        m = (y[2] - y[1]) / (x[2] - x[1]);
        intercept = y[1] - m * x[1];
    } else {
        long ncomb;
        double *mbs;
        if (complete) {
            ncomb = n * (n - 1) / 2;
            mbs = (double *)malloc(sizeof(double) * (size_t)(ncomb > n ? ncomb : n));
            long index = 0;
            for (long i = 1; i < n; i++)
                for (long j = i + 1; j <= n; j++) mbs[index++] = (y[j] - y[i]) / (x[j] - x[i]);
        } else {
            long numberOfPairs = n / 2;
            long n2 = (n % 2 == 1) ? numberOfPairs + 1 : numberOfPairs;
            ncomb = numberOfPairs;
            mbs = (double *)malloc(sizeof(double) * (size_t)n);
            for (long i = 1; i <= numberOfPairs; i++) {
                long i2 = n2 + i;
                mbs[i - 1] = (y[i2] - y[i]) / (x[i2] - x[i]);
            }
        }
        sort_doubles(mbs, ncomb);
        m = NUMquantile(mbs - 1, ncomb, 0.5);
        for (long i = 1; i <= n; i++) mbs[i - 1] = y[i] - m * x[i];
        sort_doubles(mbs, n);
        intercept = NUMquantile(mbs - 1, n, 0.5);
        free(mbs);
    }
    if (out_m) *out_m = m;
    if (out_intercept) *out_intercept = intercept;
}

/* ---------------- FFT ---------------- */

/* In-place iterative radix-2 complex FFT (n a power of two).  sign=-1: X[k]=sum x e^{-2 pi i kn/N}.
 * Stands in for Praat's FFTPACK-based NUMfft_forward/backward (dwsys/NUMfft_core.h): same DFT, other
 * operation order (rounding differs at the 1e-16 level). */
void fft_pow2(double *re, double *im, long n, int sign) {
    long j = 0;
    for (long i = 0; i < n - 1; i++) {
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
        long m = n >> 1;
        while (m >= 1 && (j & m)) { j ^= m; m >>= 1; }
        j |= m;
    }
    /* twiddle table cached per size */
    static __thread double *tw_c = NULL, *tw_s = NULL;
    static __thread long tw_n = 0;
    if (tw_n != n) {
        free(tw_c); free(tw_s);
        tw_c = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1));
        tw_s = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1));
        for (long k = 0; k < n / 2; k++) {
            tw_c[k] = cos(2.0 * NUMpi * (double)k / (double)n);
            tw_s[k] = sin(2.0 * NUMpi * (double)k / (double)n);
        }
        tw_n = n;
    }
    for (long len = 2; len <= n; len <<= 1) {
        long half = len >> 1, step = n / len;
        for (long i = 0; i < n; i += len) {
            for (long k = 0; k < half; k++) {
                double wr = tw_c[k * step], wi = sign * tw_s[k * step];
                long a = i + k, b = a + half;
                double xr = re[b] * wr - im[b] * wi;
                double xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
        }
    }
}

/* ---------------- Vector queries ---------------- */

/* fon/Vector.cpp Vector_getValueAtX (used at mshds_extractor.py:85,101 with "Cubic") */
double vector_getValueAtX(const Contour *c, double x, int interpolation) {
    double leftEdge = c->x1 - 0.5 * c->dx, rightEdge = leftEdge + c->nx * c->dx;
    if (x < leftEdge || x > rightEdge) return UNDEF;
    long depth = interpolation == 3 ? 70 : interpolation == 4 ? 700 : interpolation;
    return NUM_interpolate_sinc(c->y - 1, c->nx, (x - c->x1) / c->dx + 1.0, depth);
}

static void vector_getExtremumAndX(const Contour *c, double xmin, double xmax, int interpolation, double *out_ext,
                                   double *out_x, int isMaximum) {
    const double *y = c->y - 1;
    double ext, x;
    long imin, imax;
    if (xmax <= xmin) { xmin = c->xmin; xmax = c->xmax; }
    if (!getWindowSamples(c->x1, c->dx, c->nx, xmin, xmax, &imin, &imax)) {
        int itp = interpolation > 0 ? 1 : 0;
        double yleft = vector_getValueAtX(c, xmin, itp);
        double yright = vector_getValueAtX(c, xmax, itp);
        if (isMaximum) {
            ext = yleft > yright ? yleft : yright;
            x = yleft == yright ? (xmin + xmax) / 2 : yleft > yright ? xmin : xmax;
        } else {
            ext = yleft < yright ? yleft : yright;
            x = yleft == yright ? (xmin + xmax) / 2 : yleft < yright ? xmin : xmax;
        }
    } else {
        ext = y[imin];
        x = imin;
        if (isMaximum ? (y[imax] > ext) : (y[imax] < ext)) { ext = y[imax]; x = imax; }
        if (imin == 1) imin++;
        if (imax == c->nx) imax--;
        for (long i = imin; i <= imax; i++) {
            int isExt = isMaximum ? (y[i] > y[i - 1] && y[i] >= y[i + 1]) : (y[i] < y[i - 1] && y[i] <= y[i + 1]);
            if (isExt) {
                double i_real;
                double local = NUMimproveExtremum(y, c->nx, i, interpolation, &i_real, isMaximum);
                if (isMaximum ? (local > ext) : (local < ext)) { ext = local; x = i_real; }
            }
        }
        x = c->x1 + (x - 1) * c->dx;
        if (x < xmin) x = xmin; else if (x > xmax) x = xmax;
    }
    if (out_ext) *out_ext = ext;
    if (out_x) *out_x = x;
}

/* fon/Vector.cpp Vector_getMaximumAndX / Vector_getMinimumAndX (mshds_extractor.py:42,43,97,200,201) */
void vector_getMaximumAndX(const Contour *c, double xmin, double xmax, int interpolation, double *maximum, double *xOfMax) {
    vector_getExtremumAndX(c, xmin, xmax, interpolation, maximum, xOfMax, 1);
}
void vector_getMinimumAndX(const Contour *c, double xmin, double xmax, int interpolation, double *minimum, double *xOfMin) {
    vector_getExtremumAndX(c, xmin, xmax, interpolation, minimum, xOfMin, 0);
}

/* fon/Sampled.cpp Sampled_getQuantile over the whole domain (mshds_extractor.py:47) */
double contour_getQuantile(const Contour *c, double q) {
    double *v = (double *)malloc(sizeof(double) * (size_t)(c->nx > 0 ? c->nx : 1));
    long n = 0;
    for (long i = 0; i < c->nx; i++)
        if (isdefined(c->y[i])) v[n++] = c->y[i];
    double result = UNDEF;
    if (n >= 1) {
        sort_doubles(v, n);
        result = NUMquantile(v - 1, n, q);
    }
    free(v);
    return result;
}
