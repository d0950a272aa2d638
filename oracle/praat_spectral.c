/*
 * oracle/praat_spectral.c -- TEST INFRASTRUCTURE (see praat_core.h header).  PARITY UNPINNED.
 *
 * Restates Praat 6.1.38 fon/Ltas.cpp (PointProcess_Sound_to_Ltas, Ltas_getSlope), dwtools/Ltas_extensions.cpp
 * (Ltas_fitTiltLine), fon/Sound.cpp (Sound_resample, Sound_extractPart, Sound_preEmphasis),
 * fon/Sound_to_Formant.cpp (Sound_to_Formant_burg), dwsys/NUM2.cpp (VECburg), dwsys/Roots.cpp
 * (Polynomial_to_Roots + polish + Roots_fixIntoUnitCircle), LPC/Sound_to_PowerCepstrogram + PowerCepstrogram.cpp
 * + PowerCepstrum.cpp (CPPS), fon/Sound_and_Spectrogram.cpp + fon/Spectrum.cpp (spectral moments).
 * Serves mshds_extractor.py:241-248 (LTAS), :286-293 (CPP), :319-331 (formants), :356-369 (moments).
 */
#include "praat_core.h"
#include <complex.h>

/* ---------------- LTAS ---------------- */

/* fon/Ltas.cpp PointProcess_Sound_to_Ltas (mshds_extractor.py:241 after Sound_to_PointProcess_periodic_cc). */
int pointprocess_sound_to_ltas(const Points *pulses, const Sound *sound, double maximumFrequency, double bandWidth,
                               double shortestPeriod, double longestPeriod, double maximumPeriodFactor,
                               double *ltas /*0-based [nbands]*/, long *nbands_out) {
    long numberOfPeriods = pulses->n - 2, totalNumberOfEnergies = 0;
    long nb = (long)floor(maximumFrequency / bandWidth);
    *nbands_out = nb;
    double *numbers = (double *)calloc((size_t)nb, sizeof(double));
    for (long i = 0; i < nb; i++) ltas[i] = 0.0;
    if (numberOfPeriods < 1) { free(numbers); return 0; }
    const double *t = pulses->t - 1;    /* 1-based */
    for (long ipulse = 2; ipulse < pulses->n; ipulse++) {
        double leftInterval = t[ipulse] - t[ipulse - 1];
        double rightInterval = t[ipulse + 1] - t[ipulse];
        double intervalFactor = leftInterval > rightInterval ? leftInterval / rightInterval : rightInterval / leftInterval;
        if (leftInterval >= shortestPeriod && leftInterval <= longestPeriod && rightInterval >= shortestPeriod &&
            rightInterval <= longestPeriod && intervalFactor <= maximumPeriodFactor) {
            double ta = t[ipulse] - 0.5 * leftInterval, tb = t[ipulse] + 0.5 * rightInterval;
            /* Sound_extractPart rectangular: samples ix1..ix2, virtual samples outside the sound are zero */
            long ix1 = 1 + (long)ceil((ta - sound->x1) / sound->dx);
            long ix2 = 1 + (long)floor((tb - sound->x1) / sound->dx);
            if (ix2 < ix1) { free(numbers); return 0; }    /* "Extracted Sound would contain no samples" */
            long n = ix2 - ix1 + 1;
            /* Sound_to_Spectrum (fast = false): exact-length DFT, scaled by dx; spectrum dx = 1/(dx*n) */
            long nfreq = n / 2 + 1;
            double df = 1.0 / (sound->dx * n);
            double *ctab = (double *)malloc(sizeof(double) * (size_t)n * 3), *stab = ctab + n, *seg = stab + n;
            for (long j = 0; j < n; j++) {
                double ang = 2.0 * NUMpi * (double)j / (double)n;
                ctab[j] = cos(ang); stab[j] = sin(ang);
                long is = ix1 + j;
                seg[j] = (is >= 1 && is <= sound->nx) ? Z(sound, is) : 0.0;
            }
            for (long ifreq = 1; ifreq <= nfreq; ifreq++) {
                double frequency = (ifreq - 1) * df;
                long iband = (long)ceil(frequency / bandWidth);
                if (!(iband >= 1 && iband <= nb)) continue;
                double re = 0.0, im = 0.0;
                long ph = 0;
                for (long j = 0; j < n; j++) {
                    /* phase index (ifreq-1)*j reduced modulo n for accuracy */
                    re += seg[j] * ctab[ph];
                    im -= seg[j] * stab[ph];
                    ph += ifreq - 1;
                    if (ph >= n) ph -= n;
                }
                re *= sound->dx; im *= sound->dx;
                if (ifreq == nfreq && (n & 1) == 0) im = 0.0;   /* Nyquist bin of an even-length spectrum */
                double energy = (re * re + im * im) * 2.0 * df;
                ltas[iband - 1] += energy;
                numbers[iband - 1] += 1;
                totalNumberOfEnergies += 1;
            }
            free(ctab);
        } else {
            numberOfPeriods -= 1;
        }
    }
    if (numberOfPeriods < 1) { free(numbers); return 0; }
    for (long iband = 0; iband < nb; iband++) {
        if (numbers[iband] == 0.0) {
            ltas[iband] = UNDEF;
        } else {
            double totalEnergyInThisBand = ltas[iband];
            double meanEnergyInThisBand = totalEnergyInThisBand / numbers[iband];
            double meanNumberOfEnergiesPerBand = (double)totalNumberOfEnergies / nb;
            double redistributedEnergyInThisBand = meanEnergyInThisBand * meanNumberOfEnergiesPerBand;
            double redistributedEnergyDensityInThisBand = redistributedEnergyInThisBand / bandWidth;
            double redistributedPowerDensityInThisBand = redistributedEnergyDensityInThisBand / (sound->xmax - sound->xmin);
            ltas[iband] = 10.0 * log10(redistributedPowerDensityInThisBand / 4.0e-10);
        }
    }
    double x1 = 0.5 * bandWidth;
    double *measured = NULL;
    if (orc_opt.ltas_fill == 1) {      /* alternative: interpolate between MEASURED bands only */
        measured = (double *)malloc(sizeof(double) * (size_t)nb);
        memcpy(measured, ltas, sizeof(double) * (size_t)nb);
    }
    for (long iband = 1; iband <= nb; iband++) {
        if (measured ? isundef(measured[iband - 1]) : isundef(ltas[iband - 1])) {
            long ibandleft = iband - 1, ibandright = iband + 1;
            if (measured) {
                while (ibandleft >= 1 && isundef(measured[ibandleft - 1])) ibandleft--;
                while (ibandright <= nb && isundef(measured[ibandright - 1])) ibandright++;
                if (ibandleft < 1 && ibandright > nb) { free(numbers); free(measured); return 0; }
                if (ibandleft < 1) ltas[iband - 1] = measured[ibandright - 1];
                else if (ibandright > nb) ltas[iband - 1] = measured[ibandleft - 1];
                else {
                    double frequency = x1 + (iband - 1) * bandWidth;
                    double fleft = x1 + (ibandleft - 1) * bandWidth;
                    double fright = x1 + (ibandright - 1) * bandWidth;
                    ltas[iband - 1] = ((fright - frequency) * measured[ibandleft - 1] + (frequency - fleft) * measured[ibandright - 1]) / (fright - fleft);
                }
                continue;
            }
            while (ibandleft >= 1 && isundef(ltas[ibandleft - 1])) ibandleft--;
            while (ibandright <= nb && isundef(ltas[ibandright - 1])) ibandright++;
            if (ibandleft < 1 && ibandright > nb) { free(numbers); return 0; }
            if (ibandleft < 1) ltas[iband - 1] = ltas[ibandright - 1];
            else if (ibandright > nb) ltas[iband - 1] = ltas[ibandleft - 1];
            else {
                double frequency = x1 + (iband - 1) * bandWidth;
                double fleft = x1 + (ibandleft - 1) * bandWidth;
                double fright = x1 + (ibandright - 1) * bandWidth;
                ltas[iband - 1] = ((fright - frequency) * ltas[ibandleft - 1] + (frequency - fleft) * ltas[ibandright - 1]) / (fright - fleft);
            }
        }
    }
    free(numbers);
    free(measured);
    return 1;
}

/* fon/Sampled.cpp Sampled_getMean, interpolate = false, on the Ltas (x1 = dx/2) in dB units. */
static double ltas_mean_dB(const double *z0, long nx, double dx, double xmin, double xmax) {
    double x1 = 0.5 * dx;
    long double sum = 0.0L, definitionRange = 0.0L;
    double dom_min = 0.0, dom_max = nx * dx;
    if (xmax <= xmin) { xmin = dom_min; xmax = dom_max; }   /* Function_unidirectionalAutowindow */
    if (xmin < dom_min) xmin = dom_min;
    if (xmax > dom_max) xmax = dom_max;
    if (xmin >= xmax) return UNDEF;
    double rimin = (xmin - x1) / dx + 1.0, rimax = (xmax - x1) / dx + 1.0;
    if (rimax >= 0.5 && rimin < nx + 0.5) {
        long imin = rimin < 0.5 ? 0 : iround(rimin);
        long imax = rimax >= nx + 0.5 ? nx + 1 : iround(rimax);
        for (long isamp = imin + 1; isamp < imax; isamp++) {
            double value = z0[isamp - 1];
            if (isdefined(value)) { definitionRange += 1.0; sum += value; }
        }
        if (imin == imax) {
            if (imin >= 1 && imin <= nx && isdefined(z0[imin - 1])) {
                double phase = rimax - rimin;
                definitionRange += phase; sum += phase * z0[imin - 1];
            }
        } else {
            if (imin >= 1 && isdefined(z0[imin - 1])) {
                double phase = imin - rimin + 0.5;
                definitionRange += phase; sum += phase * z0[imin - 1];
            }
            if (imax <= nx && isdefined(z0[imax - 1])) {
                double phase = rimax - imax + 0.5;
                definitionRange += phase; sum += phase * z0[imax - 1];
            }
        }
    }
    if (definitionRange <= 0.0L) return UNDEF;
    return (double)(sum / definitionRange);
}

/* fon/Ltas.cpp Ltas_getSlope with averagingUnits = dB (mshds_extractor.py:242). */
double ltas_getSlope_dB(const double *z0, long nx, double dx, double f1min, double f1max, double f2min, double f2max) {
    double low = ltas_mean_dB(z0, nx, dx, f1min, f1max);
    double high = ltas_mean_dB(z0, nx, dx, f2min, f2max);
    if (isundef(low) || isundef(high)) return UNDEF;
    return high - low;
}

/* dwtools/Ltas_extensions.cpp Ltas_fitTiltLine (Linear frequency scale, Robust = incomplete Theil)
 * behind "Report spectral tilt 100 5000 Linear Robust" (mshds_extractor.py:245-248; the report text prints the
 * slope with round-trip precision so float() of it is the value itself). */
int ltas_fitTiltLine_robust(const double *z0, long nx, double dx, double fmin, double fmax, double *slope, double *intercept) {
    double x1 = 0.5 * dx;
    if (fmax <= fmin) { fmin = 0.0; fmax = nx * dx; }
    long ifmin, ifmax;
    long n = getWindowSamples(x1, dx, nx, fmin, fmax, &ifmin, &ifmax);
    if (n < 2) return 0;
    double *x = (double *)malloc(sizeof(double) * (size_t)n);
    double *y = (double *)malloc(sizeof(double) * (size_t)n);
    for (long i = ifmin; i <= ifmax; i++) {
        x[i - ifmin] = x1 + (i - 1) * dx;
        y[i - ifmin] = z0[i - 1];
    }
    NUMlineFit_theil(x - 1, y - 1, n, slope, intercept, orc_opt.theil_tilt_complete);
    free(x); free(y);
    return 1;
}

/* ---------------- Sound manipulation ---------------- */

/* fon/Sound.cpp Sound_resample (mshds_extractor.py:419 prec 50; inside To Formant (burg) prec 500; inside
 * To PowerCepstrogram prec 50). */
/* fon/Sound.cpp Sound_upsample: exact doubling of the sampling frequency.  nfft = power of two >= nx + 2000, the sound sits
 * at data[1001 .. 1000 + nx] of a zeroed buffer of 2 nfft reals; NUMrealft forward over the first nfft; the packed values
 * data[i], i > imin = (long)(0.95 nfft), are tapered by (nfft - i) / (nfft - imin) (real and imaginary part of a bin carry
 * consecutive i), data[2] (Nyquist) = 0; NUMrealft inverse over all 2 nfft; new sample i = data[i + 2000] / nfft.
 * New time domain: nx' = 2 nx, dx' = dx / 2, x1' = x1 - dx / 4. */
Sound *sound_upsample(const Sound *me) {
    long nfft = 1;
    while (nfft < me->nx + 2000) nfft *= 2;
    Sound *thee = sound_create(me->xmin, me->xmax, me->nx * 2, me->dx / 2.0, me->x1 - me->dx / 4.0);
    double *re = (double *)calloc((size_t)2 * nfft, sizeof(double));
    double *im = (double *)calloc((size_t)2 * nfft, sizeof(double));
    for (long i = 1; i <= me->nx; i++) re[1000 + i - 1] = Z(me, i);
    fft_pow2(re, im, nfft, -1);
    /* packed layout of the half spectrum: data[1] = DC, data[2] = Nyquist, data[2k+1] = Re X_k, data[2k+2] = Im X_k */
    long imin = (long)(nfft * 0.95);
    double *sre = (double *)calloc((size_t)2 * nfft, sizeof(double));
    double *sim = (double *)calloc((size_t)2 * nfft, sizeof(double));
    sre[0] = re[0];                                            /* data[1]: never tapered (imin >= 1) */
    for (long k = 1; k < nfft / 2; k++) {
        long ir = 2 * k + 1, ii = 2 * k + 2;
        double fr = ir > imin ? (double)(nfft - ir) / (double)(nfft - imin) : 1.0;
        double fi = ii > imin ? (double)(nfft - ii) / (double)(nfft - imin) : 1.0;
        sre[k] = re[k] * fr; sim[k] = im[k] * fi;
        sre[2 * nfft - k] = sre[k]; sim[2 * nfft - k] = -sim[k];      /* Hermitian completion of the 2 nfft spectrum */
    }
    fft_pow2(sre, sim, 2 * nfft, +1);
    double factor = 1.0 / nfft;
    for (long i = 1; i <= thee->nx; i++) Z(thee, i) = sre[i + 2000 - 1] * factor;
    free(re); free(im); free(sre); free(sim);
    return thee;
}

Sound *sound_resample(const Sound *me, double samplingFrequency, long precision) {
    double upfactor = samplingFrequency * me->dx;
    if (fabs(upfactor - 2) < 1e-6) return sound_upsample(me);
    if (fabs(upfactor - 1) < 1e-6) return sound_copy(me);
    long numberOfSamples = iround((me->xmax - me->xmin) * samplingFrequency);
    if (numberOfSamples < 1) return NULL;
    Sound *filtered = NULL;
    const Sound *src = me;
    if (upfactor < 1.0) {
        long nfft = 1, antiTurnAround = 1000;
        while (nfft < me->nx + antiTurnAround * 2) nfft *= 2;
        double *re = (double *)calloc((size_t)nfft, sizeof(double));
        double *im = (double *)calloc((size_t)nfft, sizeof(double));
        filtered = sound_create(me->xmin, me->xmax, me->nx, me->dx, me->x1);
        /* data[antiTurnAround + i] = z[i], i = 1..nx (1-based data) */
        for (long i = 1; i <= me->nx; i++) re[antiTurnAround + i - 1] = Z(me, i);
        fft_pow2(re, im, nfft, -1);
        /* NUMrealft packed layout: data[1]=DC, data[2]=Nyquist, data[2k+1]=Re X_k, data[2k+2]=Im X_k.
         * Praat zeroes data[i] for i = floor(upfactor*nfft) .. nfft and data[2]. */
        long i0 = (long)floor(upfactor * nfft);
        for (long i = i0; i <= nfft; i++) {
            if (i < 3) continue;
            long k = (i - 1) / 2;
            if (i & 1) re[k] = 0.0; else im[k] = 0.0;
        }
        re[nfft / 2] = 0.0; im[nfft / 2] = 0.0;    /* Nyquist */
        if (i0 <= 1) re[0] = 0.0;
        /* restore Hermitian symmetry for the complex inverse */
        for (long k = 1; k < nfft / 2; k++) { re[nfft - k] = re[k]; im[nfft - k] = -im[k]; }
        im[0] = 0.0;
        fft_pow2(re, im, nfft, +1);
        double factor = 1.0 / nfft;
        for (long i = 1; i <= me->nx; i++) Z(filtered, i) = re[i + antiTurnAround - 1] * factor;
        free(re); free(im);
        src = filtered;
    }
    Sound *thee = sound_create(me->xmin, me->xmax, numberOfSamples, 1.0 / samplingFrequency,
                               0.5 * (me->xmin + me->xmax - (numberOfSamples - 1) / samplingFrequency));
    if (precision <= 1) {
        for (long i = 1; i <= numberOfSamples; i++) {
            double x = s_indexToX(thee, i);
            double index = s_xToIndex(src, x);
            long leftSample = (long)floor(index);
            double fraction = index - leftSample;
            Z(thee, i) = leftSample < 1 || leftSample >= src->nx ? 0.0 : (1 - fraction) * Z(src, leftSample) + fraction * Z(src, leftSample + 1);
        }
    } else {
        for (long i = 1; i <= numberOfSamples; i++) {
            double x = s_indexToX(thee, i);
            double index = s_xToIndex(src, x);
            Z(thee, i) = NUM_interpolate_sinc(src->z - 1, src->nx, index, precision);
        }
    }
    sound_free(filtered);
    return thee;
}

/* fon/Sound.cpp Sound_extractPart (rectangular, relativeWidth 1, preserveTimes false) (mshds_extractor.py:286). */
Sound *sound_extractPart(const Sound *me, double t1, double t2) {
    if (t1 == t2) { t1 = me->xmin; t2 = me->xmax; }
    long ix1 = 1 + (long)ceil((t1 - me->x1) / me->dx);
    long ix2 = 1 + (long)floor((t2 - me->x1) / me->dx);
    if (ix2 < ix1) return NULL;
    Sound *thee = sound_create(t1, t2, ix2 - ix1 + 1, me->dx, me->x1 + (ix1 - 1) * me->dx);
    thee->xmin = 0.0; thee->xmax -= t1; thee->x1 -= t1;
    long lo = ix1 < 1 ? 1 : ix1, hi = ix2 > me->nx ? me->nx : ix2;
    for (long i = lo; i <= hi; i++) Z(thee, i - ix1 + 1) = Z(me, i);
    return thee;
}

/* fon/Sound.cpp Sound_preEmphasis */
void sound_preEmphasis(Sound *me, double preEmphasisFrequency) {
    if (preEmphasisFrequency >= 0.5 / me->dx) return;
    double emphasisFactor = exp(-2.0 * NUMpi * preEmphasisFrequency * me->dx);
    for (long i = me->nx; i >= 2; i--) Z(me, i) -= emphasisFactor * Z(me, i - 1);
}

/* ---------------- Formant (Burg) ---------------- */

void formant_free(Formant *f) {
    if (!f) return;
    free(f->nFormants); free(f->f); free(f->bw);
    free(f);
}

/* dwsys/NUM2.cpp VECburg (Numerical Recipes memcof); a[1..m], x[1..n] */
double VECburg(double *a, int m, const double *x, long n) {
    for (int j = 1; j <= m; j++) a[j] = 0.0;
    if (n <= 2) { a[1] = -1.0; return n == 2 ? 0.5 * (x[1] * x[1] + x[2] * x[2]) : x[1] * x[1]; }
    double *b1 = (double *)calloc((size_t)n + 1, sizeof(double));
    double *b2 = (double *)calloc((size_t)n + 1, sizeof(double));
    double *aa = (double *)calloc((size_t)m + 1, sizeof(double));
    long double p = 0.0L;
    for (long j = 1; j <= n; j++) p += x[j] * x[j];
    long double xms = p / n;
    if (xms <= 0.0L) { free(b1); free(b2); free(aa); return (double)xms; }
    b1[1] = x[1];
    b2[n - 1] = x[n];
    for (long j = 2; j <= n - 1; j++) b1[j] = b2[j - 1] = x[j];
    for (int i = 1; i <= m; i++) {
        long double num = 0.0L, denum = 0.0L;
        for (long j = 1; j <= n - i; j++) {
            num += b1[j] * b2[j];
            denum += b1[j] * b1[j] + b2[j] * b2[j];
        }
        if (denum <= 0.0L) { free(b1); free(b2); free(aa); return 0.0; }
        a[i] = 2.0 * (double)(num / denum);
        xms *= 1.0 - a[i] * a[i];
        for (int j = 1; j <= i - 1; j++) a[j] = aa[j] - a[i] * aa[i - j];
        if (i < m) {
            for (int j = 1; j <= i; j++) aa[j] = a[j];
            for (long j = 1; j <= n - i - 1; j++) {
                b1[j] -= aa[i] * b2[j];
                b2[j] = b2[j + 1] - aa[i] * b1[j + 1];
            }
        }
    }
    free(b1); free(b2); free(aa);
    return (double)xms;
}

/* dwsys/Roots.cpp Polynomial_to_Roots: eigenvalues of the companion matrix (Praat: LAPACK dhseqr on the upper
 * Hessenberg companion) followed by Roots_Polynomial_polish (Newton-Raphson on the original polynomial).
 * Restated with a shifted-QR (Francis) iteration on the companion matrix; c[0..n] ascending powers. */
static void poly_eval_d(const double *c, int n, double complex z, double complex *p, double complex *dp) {
    double complex pv = c[n], dv = 0.0;
    for (int i = n - 1; i >= 0; i--) { dv = dv * z + pv; pv = pv * z + c[i]; }
    *p = pv; *dp = dv;
}

/* Hessenberg QR (EISPACK hqr, as in Numerical Recipes) on a[0..n-1][0..n-1] */
static int hqr(double *a, int n, double *wr, double *wi) {
#define A(i, j) a[(i) * n + (j)]
    int nn, m, l, k, j, its, i, mmin;
    double z, y, x, w, v, u, t, s, r = 0, q = 0, p = 0, anorm = 0.0;
    for (i = 0; i < n; i++)
        for (j = (i - 1 > 0 ? i - 1 : 0); j < n; j++) anorm += fabs(A(i, j));
    nn = n - 1;
    t = 0.0;
    while (nn >= 0) {
        its = 0;
        do {
            for (l = nn; l >= 1; l--) {
                s = fabs(A(l - 1, l - 1)) + fabs(A(l, l));
                if (s == 0.0) s = anorm;
                if (fabs(A(l, l - 1)) + s == s) { A(l, l - 1) = 0.0; break; }
            }
            x = A(nn, nn);
            if (l == nn) {
                wr[nn] = x + t; wi[nn--] = 0.0;
            } else {
                y = A(nn - 1, nn - 1);
                w = A(nn, nn - 1) * A(nn - 1, nn);
                if (l == nn - 1) {
                    p = 0.5 * (y - x);
                    q = p * p + w;
                    z = sqrt(fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + (p >= 0 ? fabs(z) : -fabs(z));
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = -(wi[nn] = z);
                    }
                    nn -= 2;
                } else {
                    if (its == 120) return 0;
                    if (its == 10 || its == 20 || its == 40 || its == 80) {
                        t += x;
                        for (i = 0; i <= nn; i++) A(i, i) -= x;
                        s = fabs(A(nn, nn - 1)) + fabs(A(nn - 1, nn - 2));
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    for (m = nn - 2; m >= l; m--) {
                        z = A(m, m);
                        r = x - z;
                        s = y - z;
                        p = (r * s - w) / A(m + 1, m) + A(m, m + 1);
                        q = A(m + 1, m + 1) - z - r - s;
                        r = A(m + 2, m + 1);
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        u = fabs(A(m, m - 1)) * (fabs(q) + fabs(r));
                        v = fabs(p) * (fabs(A(m - 1, m - 1)) + fabs(z) + fabs(A(m + 1, m + 1)));
                        if (u + v == v) break;
                    }
                    for (i = m + 2; i <= nn; i++) {
                        A(i, i - 2) = 0.0;
                        if (i != m + 2) A(i, i - 3) = 0.0;
                    }
                    for (k = m; k <= nn - 1; k++) {
                        if (k != m) {
                            p = A(k, k - 1);
                            q = A(k + 1, k - 1);
                            r = 0.0;
                            if (k != nn - 1) r = A(k + 2, k - 1);
                            if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        double sg = sqrt(p * p + q * q + r * r);
                        s = p >= 0 ? sg : -sg;
                        if (s != 0.0) {
                            if (k == m) {
                                if (l != m) A(k, k - 1) = -A(k, k - 1);
                            } else {
                                A(k, k - 1) = -s * x;
                            }
                            p += s;
                            x = p / s; y = q / s; z = r / s;
                            q /= p; r /= p;
                            for (j = k; j <= nn; j++) {
                                p = A(k, j) + q * A(k + 1, j);
                                if (k != nn - 1) { p += r * A(k + 2, j); A(k + 2, j) -= p * z; }
                                A(k + 1, j) -= p * y;
                                A(k, j) -= p * x;
                            }
                            mmin = nn < k + 3 ? nn : k + 3;
                            for (i = l; i <= mmin; i++) {
                                p = x * A(i, k) + y * A(i, k + 1);
                                if (k != nn - 1) { p += z * A(i, k + 2); A(i, k + 2) -= p * r; }
                                A(i, k + 1) -= p * q;
                                A(i, k) -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return 1;
#undef A
}

int polynomial_roots(const double *c, int n, double *re, double *im) {
    while (n > 0 && c[n] == 0.0) n--;
    if (n < 1) return 0;
    double *h = (double *)calloc((size_t)n * n, sizeof(double));
    /* upper Hessenberg companion: first row = -c[n-1-j]/c[n], subdiagonal ones */
    for (int j = 0; j < n; j++) h[j] = -c[n - 1 - j] / c[n];
    for (int i = 1; i < n; i++) h[i * n + (i - 1)] = 1.0;
    int ok = hqr(h, n, re, im);
    free(h);
    if (!ok) return 0;
    /* Roots_Polynomial_polish: Newton-Raphson, keep the iterate with the smallest |p| */
    for (int i = 0; i < n; i++) {
        double complex z = re[i] + I * im[i], p, dp;
        poly_eval_d(c, n, z, &p, &dp);
        double best = cabs(p);
        double complex zbest = z;
        for (int it = 0; it < 80; it++) {
            if (cabs(dp) == 0.0) break;
            double complex znew = z - p / dp;
            poly_eval_d(c, n, znew, &p, &dp);
            double fabsp = cabs(p);
            if (fabsp >= best) break;
            best = fabsp; zbest = znew; z = znew;
        }
        re[i] = creal(zbest); im[i] = cimag(zbest);
    }
    return n;
}

/* fon/Sound_to_Formant.cpp Sound_to_Formant_burg -> Sound_to_Formant_any -> _inplace + burg() (mshds_extractor.py:319). */
Formant *sound_to_formant_burg(const Sound *me_in, double dt_in, double nFormants, double maximumFrequency,
                               double halfdt_window, double preemphasisFrequency) {
    int numberOfPoles = (int)floor(2 * nFormants);
    double safetyMargin = 50.0;
    double nyquist0 = 0.5 / me_in->dx;
    Sound *me;
    if (maximumFrequency <= 0.0 || fabs(maximumFrequency / nyquist0 - 1) < 1.0e-12) me = sound_copy(me_in);
    else me = sound_resample(me_in, maximumFrequency * 2, 500);
    if (!me) return NULL;
    double dt = dt_in > 0.0 ? dt_in : halfdt_window / 4.0;
    double duration = me->nx * me->dx, t1;
    double dt_window = 2.0 * halfdt_window;
    long nFrames = 1 + (long)floor((duration - dt_window) / dt);
    long nsamp_window = (long)floor(dt_window / me->dx), halfnsamp_window = nsamp_window / 2;
    if (nsamp_window < numberOfPoles + 1) { sound_free(me); return NULL; }
    t1 = me->x1 + 0.5 * (duration - me->dx - (nFrames - 1) * dt);
    if (nFrames < 1) {
        nFrames = 1;
        t1 = me->x1 + 0.5 * duration;
        dt_window = duration;
        nsamp_window = me->nx;
        halfnsamp_window = nsamp_window / 2;
    }
    Formant *thee = (Formant *)calloc(1, sizeof(Formant));
    thee->xmin = me->xmin; thee->xmax = me->xmax; thee->nx = nFrames; thee->dx = dt; thee->x1 = t1;
    thee->maxnFormants = (numberOfPoles + 1) / 2;
    thee->nFormants = (int *)calloc((size_t)nFrames, sizeof(int));
    thee->f = (double *)calloc((size_t)nFrames * thee->maxnFormants, sizeof(double));
    thee->bw = (double *)calloc((size_t)nFrames * thee->maxnFormants, sizeof(double));
    double *window = (double *)calloc((size_t)nsamp_window + 1, sizeof(double));
    double *frame = (double *)calloc((size_t)nsamp_window + 1, sizeof(double));
    double *cof = (double *)calloc((size_t)numberOfPoles + 1, sizeof(double));
    double *poly = (double *)calloc((size_t)numberOfPoles + 1, sizeof(double));
    double *rre = (double *)calloc((size_t)numberOfPoles, sizeof(double));
    double *rim = (double *)calloc((size_t)numberOfPoles, sizeof(double));

    sound_preEmphasis(me, preemphasisFrequency);
    for (long i = 1; i <= nsamp_window; i++) {
        double imid = 0.5 * (nsamp_window + 1), edge = exp(-12.0);
        window[i] = (exp(-48.0 * (i - imid) * (i - imid) / (nsamp_window + 1) / (nsamp_window + 1)) - edge) / (1.0 - edge);
    }
    double nyquistFrequency = 0.5 / me->dx;
    for (long iframe = 1; iframe <= nFrames; iframe++) {
        double t = thee->x1 + (iframe - 1) * thee->dx;
        long leftSample = s_xToLowIndex(me, t);
        long rightSample = leftSample + 1;
        long startSample = rightSample - halfnsamp_window;
        long endSample = leftSample + halfnsamp_window;
        double maximumIntensity = 0.0;
        if (startSample < 1) startSample = 1;
        if (endSample > me->nx) endSample = me->nx;
        for (long i = startSample; i <= endSample; i++) {
            double value = Z(me, i);
            if (value * value > maximumIntensity) maximumIntensity = value * value;
        }
        if (maximumIntensity == 0.0) continue;
        for (long j = 1, i = startSample; j <= nsamp_window; j++, i++) frame[j] = (i <= me->nx ? Z(me, i) : 0.0) * window[j];
        VECburg(cof, numberOfPoles, frame, nsamp_window);
        /* polynomial z^n - sum cof[k] z^(n-k): coefficients[i] = -cof[n-i+1], coefficients[n+1] = 1 */
        for (int i = 1; i <= numberOfPoles; i++) poly[i - 1] = -cof[numberOfPoles - i + 1];
        poly[numberOfPoles] = 1.0;
        int nr = polynomial_roots(poly, numberOfPoles, rre, rim);
        int nf = 0;
        double *ff = thee->f + (iframe - 1) * thee->maxnFormants, *fb = thee->bw + (iframe - 1) * thee->maxnFormants;
        for (int i = 0; i < nr; i++) {
            double re = rre[i], im = rim[i];
            double a2 = re * re + im * im;
            if (a2 > 1.0) { re /= a2; im /= a2; }    /* Roots_fixIntoUnitCircle: z -> 1/conj(z) */
            if (im >= 0) {
                double f = fabs(atan2(im, re)) * nyquistFrequency / NUMpi;
                if (f >= safetyMargin && f <= nyquistFrequency - safetyMargin && nf < thee->maxnFormants) {
                    ff[nf] = f;
                    fb[nf] = -log(re * re + im * im) * nyquistFrequency / NUMpi;
                    nf++;
                }
            }
        }
        /* Formant_sort: by frequency */
        for (int i = 1; i < nf; i++)
            for (int j = i; j > 0 && ff[j] < ff[j - 1]; j--) {
                double tf = ff[j]; ff[j] = ff[j - 1]; ff[j - 1] = tf;
                tf = fb[j]; fb[j] = fb[j - 1]; fb[j - 1] = tf;
            }
        thee->nFormants[iframe - 1] = nf;
    }
    free(window); free(frame); free(cof); free(poly); free(rre); free(rim);
    sound_free(me);
    return thee;
}

/* Formant "Get value at time" / "Get bandwidth at time" (Hertz, Linear) (mshds_extractor.py:327-330) */
double formant_getValueAtTime(const Formant *me, int iformant, double x, int bandwidth) {
    if (x < me->xmin || x > me->xmax) return UNDEF;
    double ireal = (x - me->x1) / me->dx + 1.0;
    long ileft = (long)floor(ireal), inear, ifar;
    double phase = ireal - ileft;
    if (phase < 0.5) { inear = ileft; ifar = ileft + 1; }
    else { ifar = ileft; inear = ileft + 1; phase = 1.0 - phase; }
    if (inear < 1 || inear > me->nx) return UNDEF;
    if (iformant > me->nFormants[inear - 1]) return UNDEF;
    const double *arr = bandwidth ? me->bw : me->f;
    double fnear = arr[(inear - 1) * me->maxnFormants + iformant - 1];
    if (ifar < 1 || ifar > me->nx) return fnear;
    if (iformant > me->nFormants[ifar - 1]) return fnear;
    double ffar = arr[(ifar - 1) * me->maxnFormants + iformant - 1];
    return fnear + phase * (ffar - fnear);
}

/* ---------------- PowerCepstrogram / CPPS ---------------- */

/* dwsys/NUM2 VECsmoothByMovingAverage_preallocated; 0-based here */
static void smooth_moving_average(double *out, const double *in, long n, long window) {
    for (long i = 1; i <= n; i++) {
        long jfrom = i - window / 2, jto = i + window / 2;
        if ((window % 2) == 0) { if (orc_opt.cpps_smooth_align == 1) jfrom++; else jto--; }
        jfrom = jfrom < 1 ? 1 : jfrom;
        jto = jto > n ? n : jto;
        double s = 0.0;
        for (long j = jfrom; j <= jto; j++) s += in[j - 1];
        out[i - 1] = s / (jto - jfrom + 1);
    }
}

/* LPC/Sound_and_PowerCepstrogram.cpp Sound_to_PowerCepstrogram + PowerCepstrogram_getCPPS
 * (mshds_extractor.py:289,291: 'To PowerCepstrogram', 60, 0.002, 5000, 50; 'Get CPPS...', no, 0.01, 0.001, 60, 330,
 * 0.05, parabolic, 0.001, 0, Straight, Robust). */
int sound_cpps(const Sound *me, double pitchFloor, double dt, double maximumFrequency, double preEmphasisFrequency,
               double timeAveragingWindow, double quefrencyAveragingWindow, double peakFloor, double peakCeiling,
               double qstartFit, double qendFit, double *cpps) {
    double analysisWidth = 3.0 / pitchFloor;
    double windowDuration = 2.0 * analysisWidth;
    long nFrames;
    if (windowDuration > me->dx * me->nx) windowDuration = me->dx * me->nx;
    double t1, samplingFrequency = 2.0 * maximumFrequency;
    Sound *sound = sound_resample(me, samplingFrequency, 50);
    if (!sound) return 0;
    sound_preEmphasis(sound, preEmphasisFrequency);
    if (!shortTermAnalysis(me->nx, me->dx, me->x1, windowDuration, dt, &nFrames, &t1)) { sound_free(sound); return 0; }
    /* Sound_createSimple (1, windowDuration, samplingFrequency): nx = floor(duration*fs + 0.5) */
    long nwin = (long)floor(windowDuration * samplingFrequency + 0.5);
    if (nwin < 1) { sound_free(sound); return 0; }
    double *window = (double *)calloc((size_t)nwin, sizeof(double));
    {
        double imid = 0.5 * (nwin + 1), edge = exp(-12.0);
        for (long i = 1; i <= nwin; i++)
            window[i - 1] = (exp(-48.0 * (i - imid) * (i - imid) / (nwin + 1) / (nwin + 1)) - edge) / (1.0 - edge);
    }
    long nfft = 2;
    while (nfft < nwin) nfft *= 2;
    long nq = nfft / 2 + 1;
    double qmax = 0.5 * nfft / samplingFrequency, dq = qmax / (nq - 1);
    double *z = (double *)calloc((size_t)nq * nFrames, sizeof(double));    /* z[iq*nFrames + iframe] */
    double *re = (double *)calloc((size_t)nfft, sizeof(double));
    double *im = (double *)calloc((size_t)nfft, sizeof(double));
    double sdx = sound->dx;
    for (long iframe = 1; iframe <= nFrames; iframe++) {
        double t = t1 + (iframe - 1) * dt;
        /* Sound_into_Sound */
        long index = s_xToNearestIndex(sound, t - windowDuration / 2);
        double mean = 0.0;
        for (long i = 1; i <= nwin; i++) {
            long j = index - 1 + i;
            re[i - 1] = (j < 1 || j > sound->nx) ? 0.0 : Z(sound, j);
            mean += re[i - 1];
        }
        mean /= nwin;     /* Vector_subtractMean */
        for (long i = 0; i < nwin; i++) { re[i] = (re[i] - mean) * window[i]; im[i] = 0.0; }
        for (long i = nwin; i < nfft; i++) { re[i] = 0.0; im[i] = 0.0; }
        fft_pow2(re, im, nfft, -1);
        /* Sound_to_Spectrum scaling dx; Spectrum_to_PowerCepstrum: log(|X|^2 + 1e-300) then Spectrum_to_Sound (scaling df) */
        double df = 1.0 / (sdx * nfft);
        for (long k = 0; k <= nfft / 2; k++) {
            double xr = re[k] * sdx, xi = im[k] * sdx;
            if (k == 0 || k == nfft / 2) xi = 0.0;
            double L = log(xr * xr + xi * xi + 1e-300);
            re[k] = L * df; im[k] = 0.0;
        }
        for (long k = 1; k < nfft / 2; k++) { re[nfft - k] = re[k]; im[nfft - k] = 0.0; }
        fft_pow2(re, im, nfft, +1);
        for (long i = 0; i < nq; i++) z[i * nFrames + (iframe - 1)] = re[i] * re[i];
    }
    free(re); free(im); free(window);
    sound_free(sound);

    /* PowerCepstrogram_smooth */
    long numberOfFrames = (long)floor(timeAveragingWindow / dt);
    if (orc_opt.cpps_time_frames == 1) numberOfFrames -= 1;
    if (numberOfFrames > 1) {
        double *qout = (double *)malloc(sizeof(double) * (size_t)nFrames);
        for (long iq = 0; iq < nq; iq++) {
            smooth_moving_average(qout, z + iq * nFrames, nFrames, numberOfFrames);
            memcpy(z + iq * nFrames, qout, sizeof(double) * (size_t)nFrames);
        }
        free(qout);
    }
    long numberOfQuefrencyBins = (long)floor(quefrencyAveragingWindow / dq);
    double *col = (double *)malloc(sizeof(double) * (size_t)nq);
    double *col2 = (double *)malloc(sizeof(double) * (size_t)nq);
    double *xq = (double *)malloc(sizeof(double) * (size_t)nq);
    long double sum = 0.0L;
    Contour c;
    c.xmin = 0.0; c.xmax = qmax; c.nx = nq; c.dx = dq; c.x1 = 0.0; c.y = col2;
    for (long iframe = 0; iframe < nFrames; iframe++) {
        for (long iq = 0; iq < nq; iq++) col[iq] = z[iq * nFrames + iframe];
        if (numberOfQuefrencyBins > 1) smooth_moving_average(col2, col, nq, numberOfQuefrencyBins);
        else memcpy(col2, col, sizeof(double) * (size_t)nq);
        /* PowerCepstrum in dB */
        for (long iq = 0; iq < nq; iq++) col2[iq] = 10.0 * log10(col2[iq] + 1e-30);
        /* PowerCepstrum_fitTiltLine: Straight, Robust; qmax <= qmin -> whole domain */
        double qlo = qstartFit, qhi = qendFit;
        if (qhi <= qlo) { qlo = orc_opt.cpps_fit_range == 1 ? qstartFit : 0.0; qhi = qmax; }
        long imin, imax;
        if (!getWindowSamples(0.0, dq, nq, qlo, qhi, &imin, &imax) || imax - imin + 1 < 2) { sum = NAN; break; }
        long npts = imax - imin + 1;
        for (long i = 0; i < npts; i++) xq[i] = (imin + i - 1) * dq;
        double slope, intercept;
        NUMlineFit_theil(xq - 1, col2 + (imin - 1) - 1, npts, &slope, &intercept, orc_opt.theil_cpps_complete);
        double peakdB, quefrency;
        vector_getMaximumAndX(&c, 1.0 / peakCeiling, 1.0 / peakFloor, PEAK_PARABOLIC, &peakdB, &quefrency);
        sum += peakdB - (slope * quefrency + intercept);
    }
    *cpps = (double)(sum / nFrames);
    free(col); free(col2); free(xq); free(z);
    return 1;
}

/* ---------------- Spectrogram moments ---------------- */

/* fon/Sound_and_Spectrogram.cpp Sound_to_Spectrogram (Gaussian), fon/Spectrogram.cpp Spectrogram_to_Spectrum,
 * fon/Spectrum.cpp central moments with power = 2 (mshds_extractor.py:356-374). */
int sound_spectral_moments(const Sound *me, const Pitch *pitch, double effectiveAnalysisWidth, double fmax,
                           double minimumTimeStep1, double minimumFreqStep1, double out4[4]) {
    double nyquist = 0.5 / me->dx;
    double physicalAnalysisWidth = 2 * effectiveAnalysisWidth;
    double effectiveTimeWidth = effectiveAnalysisWidth / sqrt(NUMpi);
    double effectiveFreqWidth = 1 / effectiveTimeWidth;
    double minimumTimeStep2 = effectiveTimeWidth / 8.0;
    double minimumFreqStep2 = effectiveFreqWidth / 8.0;
    double timeStep = minimumTimeStep1 > minimumTimeStep2 ? minimumTimeStep1 : minimumTimeStep2;
    double freqStep = minimumFreqStep1 > minimumFreqStep2 ? minimumFreqStep1 : minimumFreqStep2;
    double duration = me->dx * (double)me->nx, windowssq = 0.0;
    long nsamp_window = (long)floor(physicalAnalysisWidth / me->dx);
    long halfnsamp_window = nsamp_window / 2 - 1;
    nsamp_window = halfnsamp_window * 2;
    if (nsamp_window < 1) return 0;
    if (physicalAnalysisWidth > duration) return 0;
    long numberOfTimes = 1 + (long)floor((duration - physicalAnalysisWidth) / timeStep);
    double t1 = me->x1 + 0.5 * ((double)(me->nx - 1) * me->dx - (double)(numberOfTimes - 1) * timeStep);
    if (fmax <= 0.0 || fmax > nyquist) fmax = nyquist;
    long numberOfFreqs = (long)floor(fmax / freqStep);
    if (numberOfFreqs < 1) return 0;
    long nsampFFT = 1;
    while (nsampFFT < nsamp_window || nsampFFT < 2 * numberOfFreqs * (nyquist / fmax)) nsampFFT *= 2;
    long half_nsampFFT = nsampFFT / 2;
    long binWidth_samples = (long)floor(freqStep * me->dx * nsampFFT);
    if (binWidth_samples < 1) binWidth_samples = 1;
    double binWidth_hertz = 1.0 / (me->dx * nsampFFT);
    freqStep = binWidth_samples * binWidth_hertz;
    numberOfFreqs = (long)floor(fmax / freqStep);
    if (numberOfFreqs < 1) return 0;
    double y1 = 0.5 * (freqStep - binWidth_hertz);
    double *window = (double *)calloc((size_t)nsamp_window + 1, sizeof(double));
    for (long i = 1; i <= nsamp_window; i++) {
        double nSamplesPerWindow_f = physicalAnalysisWidth / me->dx;
        double imid = 0.5 * (double)(nsamp_window + 1), edge = exp(-12.0);
        double phase = ((double)i - imid) / nSamplesPerWindow_f;
        double value = (exp(-48.0 * phase * phase) - edge) / (1.0 - edge);
        window[i] = (float)value;
        windowssq += value * value;
    }
    double oneByBinWidth = 1.0 / windowssq / binWidth_samples;
    double *re = (double *)calloc((size_t)nsampFFT, sizeof(double));
    double *im = (double *)calloc((size_t)nsampFFT, sizeof(double));
    double *power = (double *)calloc((size_t)numberOfFreqs, sizeof(double));
    long double acc[4] = {0, 0, 0, 0};
    long cnt[4] = {0, 0, 0, 0};
    for (long iframe = 1; iframe <= numberOfTimes; iframe++) {
        double t = t1 + (iframe - 1) * timeStep;
        if (isundef(pitch_getValueAtTime(pitch, t))) continue;    /* mshds_extractor.py:364 */
        long leftSample = s_xToLowIndex(me, t), rightSample = leftSample + 1;
        long startSample = rightSample - halfnsamp_window;
        for (long j = 1, i = startSample; j <= nsamp_window; j++) { re[j - 1] = Z(me, i++) * window[j]; im[j - 1] = 0.0; }
        for (long j = nsamp_window; j < nsampFFT; j++) { re[j] = 0.0; im[j] = 0.0; }
        fft_pow2(re, im, nsampFFT, -1);
        for (long iband = 1; iband <= numberOfFreqs; iband++) {
            long leftsample = (iband - 1) * binWidth_samples + 1, rightsample = leftsample + binWidth_samples;
            double p = 0.0;
            for (long i = leftsample; i < rightsample; i++) {
                long k = i - 1;    /* spec[i] = |X_{i-1}|^2 (DC at 1) */
                p += (k == 0 || k == half_nsampFFT) ? re[k] * re[k] : re[k] * re[k] + im[k] * im[k];
            }
            power[iband - 1] = p * oneByBinWidth;
        }
        /* Spectrogram_to_Spectrum: re = sqrt(power), im = 0; Spectrum moments with power 2: energy = re*re */
        long double sumenergy = 0.0L, sumfenergy = 0.0L;
        for (long i = 1; i <= numberOfFreqs; i++) {
            double a = sqrt(power[i - 1]);
            double energy = a * a;
            double f = y1 + (i - 1) * freqStep;
            sumenergy += energy;
            sumfenergy += f * energy;
        }
        if (sumenergy == 0.0L) continue;     /* all four are undefined (NaN) and skipped (:366-369) */
        double fmean = (double)(sumfenergy / sumenergy);
        long double m2 = 0.0L, m3 = 0.0L, m4 = 0.0L;
        for (long i = 1; i <= numberOfFreqs; i++) {
            double a = sqrt(power[i - 1]);
            double energy = a * a;
            double d = y1 + (i - 1) * freqStep - fmean;
            m2 += d * d * energy;
            m3 += d * d * d * energy;
            m4 += d * d * d * d * energy;
        }
        double mu2 = (double)(m2 / sumenergy), mu3 = (double)(m3 / sumenergy), mu4 = (double)(m4 / sumenergy);
        acc[0] += fmean; cnt[0]++;
        double sd = sqrt(mu2);
        if (isdefined(sd)) { acc[1] += sd; cnt[1]++; }
        if (mu2 != 0.0) {
            acc[2] += mu3 / (mu2 * sqrt(mu2)); cnt[2]++;
            acc[3] += mu4 / (mu2 * mu2) - 3.0; cnt[3]++;
        }
    }
    for (int k = 0; k < 4; k++) out4[k] = cnt[k] > 0 ? (double)(acc[k] / cnt[k]) : UNDEF;
    free(window); free(re); free(im); free(power);
    return 1;
}
