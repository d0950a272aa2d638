/*
 * oracle/praat_intensity.c -- TEST INFRASTRUCTURE (see praat_core.h header).  PARITY UNPINNED.
 *
 * Restates Praat 6.1.38 fon/Sound_to_Intensity.cpp (Sound_to_Intensity), fon/Intensity.cpp ("Get mean ... energy"),
 * dwtools/Intensity_extensions.cpp (Intensity_to_TextGrid_detectSilences + IntervalTier_cutIntervals_minimumDuration),
 * fon/Sound_to_PointProcess.cpp (Sound_to_PointProcess_extrema), fon/Pitch_to_PointProcess.cpp
 * (Sound_Pitch_to_PointProcess_cc, Sound_findExtremum, Sound_findMaximumCorrelation).
 * Serves mshds_extractor.py:41-101 (speech rate), :198-202 (intensity), :241/:271/:321 (glottal pulses).
 */
#include "praat_core.h"

/* fon/Sound_to_Intensity.cpp Sound_to_Intensity (mshds_extractor.py:41,198); returns NULL where Praat throws. */
Contour *sound_to_intensity(const Sound *me, double minimumPitch, double timeStep, int subtractMean) {
    if (timeStep == 0.0) timeStep = 0.8 / minimumPitch;
    double physicalWindowDuration = 6.4 / minimumPitch;
    double halfWindowDuration = 0.5 * physicalWindowDuration;
    long halfWindowSamples = (long)floor(halfWindowDuration / me->dx);
    double *amplitude = (double *)calloc((size_t)(2 * halfWindowSamples + 1), sizeof(double)) + halfWindowSamples;
    double *window = (double *)calloc((size_t)(2 * halfWindowSamples + 1), sizeof(double)) + halfWindowSamples;
    for (long i = -halfWindowSamples; i <= halfWindowSamples; i++) {
        double x = i * me->dx / halfWindowDuration;
        double root = 1.0 - x * x;
        window[i] = root <= 0.0 ? 0.0 : NUMbessel_i0_f((2.0 * NUMpi * NUMpi + 0.5) * sqrt(root));
    }
    long numberOfFrames;
    double thyFirstTime;
    if (!shortTermAnalysis(me->nx, me->dx, me->x1, physicalWindowDuration, timeStep, &numberOfFrames, &thyFirstTime)) {
        free(amplitude - halfWindowSamples); free(window - halfWindowSamples);
        return NULL;
    }
    Contour *thee = (Contour *)calloc(1, sizeof(Contour));
    thee->xmin = me->xmin; thee->xmax = me->xmax; thee->nx = numberOfFrames; thee->dx = timeStep; thee->x1 = thyFirstTime;
    thee->y = (double *)calloc((size_t)numberOfFrames, sizeof(double));
    for (long iframe = 1; iframe <= numberOfFrames; iframe++) {
        double midTime = thee->x1 + (iframe - 1) * thee->dx;
        long midSample = s_xToNearestIndex(me, midTime);
        long leftSample = midSample - halfWindowSamples, rightSample = midSample + halfWindowSamples;
        long double sumxw = 0.0L, sumw = 0.0L;
        if (leftSample < 1) leftSample = 1;
        if (rightSample > me->nx) rightSample = me->nx;
        for (long i = leftSample; i <= rightSample; i++) amplitude[i - midSample] = Z(me, i);
        if (subtractMean) {
            double sum = 0.0;
            for (long i = leftSample; i <= rightSample; i++) sum += amplitude[i - midSample];
            double mean = sum / (rightSample - leftSample + 1);
            for (long i = leftSample; i <= rightSample; i++) amplitude[i - midSample] -= mean;
        }
        for (long i = leftSample; i <= rightSample; i++) {
            sumxw += amplitude[i - midSample] * amplitude[i - midSample] * window[i - midSample];
            sumw += window[i - midSample];
        }
        double intensity = (double)(sumxw / sumw);
        intensity /= 4e-10;
        thee->y[iframe - 1] = intensity < 1e-30 ? -300.0 : 10.0 * log10(intensity);
    }
    free(amplitude - halfWindowSamples); free(window - halfWindowSamples);
    return thee;
}

/* Intensity "Get mean 0 0 energy" (mshds_extractor.py:199): Sampled_getMean_standardUnit over the whole
 * domain = 10 log10(mean(10^(dB/10))). */
double intensity_getMeanEnergy(const Contour *c) {
    long double sum = 0.0L;
    long n = 0;
    for (long i = 0; i < c->nx; i++)
        if (isdefined(c->y[i])) { sum += pow(10.0, 0.1 * c->y[i]); n++; }
    if (n == 0) return UNDEF;
    return 10.0 * log10((double)(sum / n));
}

/* ---------------- silences ---------------- */

void tier_free(Tier *t) {
    if (!t) return;
    free(t->v);
    free(t);
}

/* dwtools/TextGrid_extensions.cpp IntervalTier_cutIntervals_minimumDuration; removing an interval that sits
 * between two intervals of the other label merges the three (ambiguity noted in DESIGN.md). */
static void tier_cut_short(Tier *t, int sounding, double minimumDuration) {
    long i = 0;
    while (i < t->n) {
        Interval *iv = &t->v[i];
        if (iv->sounding == sounding && iv->xmax - iv->xmin < minimumDuration && t->n > 1) {
            double xmin = iv->xmin, xmax = iv->xmax;
            if (i == 0) {
                t->v[1].xmin = xmin;
                memmove(&t->v[0], &t->v[1], sizeof(Interval) * (size_t)(t->n - 1));
                t->n -= 1;
            } else if (i == t->n - 1) {
                t->v[i - 1].xmax = xmax;
                t->n -= 1;
            } else if (orc_opt.cut_interval == 1) {
                /* alternative reading: extend the left neighbour only; i+1 (same label as i-1) stays its own interval */
                t->v[i - 1].xmax = xmax;
                memmove(&t->v[i], &t->v[i + 1], sizeof(Interval) * (size_t)(t->n - i - 1));
                t->n -= 1;
            } else {
                /* neighbours i-1 and i+1 carry the other label: merge them across the removed interval */
                t->v[i - 1].xmax = t->v[i + 1].xmax;
                memmove(&t->v[i], &t->v[i + 2], sizeof(Interval) * (size_t)(t->n - i - 2));
                t->n -= 2;
            }
            /* re-examine from the merged interval's successor: position i now holds a new interval */
        } else {
            i++;
        }
    }
}

/* dwtools/Intensity_extensions.cpp Intensity_to_TextGrid_detectSilences (mshds_extractor.py:55). */
Tier *intensity_to_silences(const Contour *me, double silenceThreshold_dB, double minSilenceDuration,
                            double minSoundingDuration) {
    Tier *t = (Tier *)calloc(1, sizeof(Tier));
    t->v = (Interval *)calloc((size_t)me->nx + 2, sizeof(Interval));
    t->n = 1;
    t->v[0].xmin = me->xmin; t->v[0].xmax = me->xmax; t->v[0].sounding = 1;
    double duration = me->xmax - me->xmin;
    if (minSilenceDuration > duration) return t;
    double intensity_max_db, intensity_min_db;
    vector_getMaximumAndX(me, 0, 0, PEAK_PARABOLIC, &intensity_max_db, NULL);
    vector_getMinimumAndX(me, 0, 0, PEAK_PARABOLIC, &intensity_min_db, NULL);
    double intensityThreshold = intensity_max_db - fabs(silenceThreshold_dB);
    if (minSilenceDuration > duration || intensityThreshold < intensity_min_db) return t;

    int inSilenceInterval = me->y[0] < intensityThreshold;
    long n = 0;
    double start = me->xmin;
    for (long i = 2; i <= me->nx; i++) {
        int silent = me->y[i - 1] < intensityThreshold;
        if (silent != inSilenceInterval) {
            double time = me->x1 + (i - 1) * me->dx;
            if (orc_opt.silence_boundary == 1) time -= 0.5 * me->dx;
            t->v[n].xmin = start; t->v[n].xmax = time; t->v[n].sounding = !inSilenceInterval;
            n++;
            start = time;
            inSilenceInterval = silent;
        }
    }
    t->v[n].xmin = start; t->v[n].xmax = me->xmax; t->v[n].sounding = !inSilenceInterval;
    t->n = n + 1;
    tier_cut_short(t, 1, minSoundingDuration);
    tier_cut_short(t, 0, minSilenceDuration);
    return t;
}

/* TextGrid "Get interval at time" (mshds_extractor.py:107): IntervalTier_timeToLowIndex */
long tier_intervalAtTime(const Tier *t, double time) {
    for (long i = 0; i < t->n; i++)
        if (time >= t->v[i].xmin && time < t->v[i].xmax) return i + 1;
    if (t->n > 0 && time == t->v[t->n - 1].xmax) return t->n;
    return 0;
}

/* ---------------- point processes ---------------- */

Points *points_create(double xmin, double xmax) {
    Points *p = (Points *)calloc(1, sizeof(Points));
    p->cap = 64; p->t = (double *)malloc(sizeof(double) * (size_t)p->cap);
    p->xmin = xmin; p->xmax = xmax;
    return p;
}
void points_free(Points *p) {
    if (!p) return;
    free(p->t);
    free(p);
}
/* fon/PointProcess.cpp PointProcess_addPoint: sorted insert */
void points_add(Points *p, double t) {
    if (p->n == p->cap) { p->cap *= 2; p->t = (double *)realloc(p->t, sizeof(double) * (size_t)p->cap); }
    long pos = p->n;
    while (pos > 0 && p->t[pos - 1] > t) pos--;
    memmove(&p->t[pos + 1], &p->t[pos], sizeof(double) * (size_t)(p->n - pos));
    p->t[pos] = t;
    p->n++;
}

/* fon/Sound_to_PointProcess.cpp Sound_to_PointProcess_extrema on the intensity contour turned Sound
 * (mshds_extractor.py:76-78: "Left", maxima yes, minima no, Sinc70). */
Points *contour_to_points_extrema_maxima(const Contour *c, int interpolation) {
    Points *p = points_create(c->xmin, c->xmax);
    const double *y = c->y - 1;
    for (long i = 2; i <= c->nx - 1; i++) {
        if (y[i] > y[i - 1] && y[i] >= y[i + 1]) {
            double i_real;
            (void)NUMimproveExtremum(y, c->nx, i, interpolation, &i_real, 1);
            points_add(p, c->x1 + (i_real - 1.0) * c->dx);
        }
    }
    return p;
}

/* fon/Pitch_to_PointProcess.cpp findExtremum_3 / Sound_findExtremum */
static double findExtremum_3(const double *channel1_base, long d, long n, int includeMaxima, int includeMinima) {
    const double *channel1 = channel1_base + d;   /* channel1[1..n] */
    int includeAll = includeMaxima == includeMinima;
    long imin = 1, imax = 1, iextr;
    double minimum, maximum;
    if (n < 3) {
        if (n <= 0) return 0.0;
        else if (n == 1) return 1.0;
        else {
            double x1 = channel1[1], x2 = channel1[2];
            double xleft = includeAll ? fabs(x1) : includeMaxima ? x1 : -x1;
            double xright = includeAll ? fabs(x2) : includeMaxima ? x2 : -x2;
            if (xleft > xright) return 1.0;
            else if (xleft < xright) return 2.0;
            else return 1.5;
        }
    }
    minimum = maximum = channel1[1];
    for (long i = 2; i <= n; i++) {
        double value = channel1[i];
        if (value < minimum) { minimum = value; imin = i; }
        if (value > maximum) { maximum = value; imax = i; }
    }
    if (minimum == maximum) return 0.5 * (n + 1.0);
    iextr = includeAll ? (fabs(minimum) > fabs(maximum) ? imin : imax) : includeMaxima ? imax : imin;
    if (iextr == 1) return 1.0;
    if (iextr == n) return (double)n;
    double valueMid = channel1[iextr], valueLeft = channel1[iextr - 1], valueRight = channel1[iextr + 1];
    return iextr + 0.5 * (valueRight - valueLeft) / (2 * valueMid - valueLeft - valueRight);
}

static double sound_findExtremum(const Sound *me, double tmin, double tmax, int includeMaxima, int includeMinima) {
    long imin = s_xToLowIndex(me, tmin), imax = s_xToHighIndex(me, tmax);
    if (imin < 1) imin = 1;
    if (imax > me->nx) imax = me->nx;
    double iextremum = findExtremum_3(me->z - 1, imin - 1, imax - imin + 1, includeMaxima, includeMinima);
    if (iextremum != 0.0) return me->x1 + (imin - 1 + iextremum - 1) * me->dx;
    return (tmin + tmax) / 2;
}

static double sound_findMaximumCorrelation(const Sound *me, double t1, double windowLength, double tmin2, double tmax2,
                                           double *tout, double *peak) {
    double maximumCorrelation = -1.0, r1 = 0.0, r2 = 0.0, r3 = 0.0, r1_best = 0.0, r3_best = 0.0, ir = 0.0;
    double halfWindowLength = 0.5 * windowLength;
    long ileft1 = s_xToNearestIndex(me, t1 - halfWindowLength);
    long iright1 = s_xToNearestIndex(me, t1 + halfWindowLength);
    long ileft2min = s_xToLowIndex(me, tmin2 - halfWindowLength);
    long ileft2max = s_xToHighIndex(me, tmax2 - halfWindowLength);
    *peak = 0.0;
    for (long ileft2 = ileft2min; ileft2 <= ileft2max; ileft2++) {
        double norm1 = 0.0, norm2 = 0.0, product = 0.0, localPeak = 0.0;
        for (long i1 = ileft1, i2 = ileft2; i1 <= iright1; i1++, i2++) {
            if (i1 < 1 || i1 > me->nx || i2 < 1 || i2 > me->nx) continue;
            double amp1 = Z(me, i1), amp2 = Z(me, i2);
            norm1 += amp1 * amp1;
            norm2 += amp2 * amp2;
            product += amp1 * amp2;
            if (fabs(amp2) > localPeak) localPeak = fabs(amp2);
        }
        r1 = r2;
        r2 = r3;
        r3 = product != 0.0 ? product / (sqrt(norm1 * norm2)) : 0.0;
        if (r2 > maximumCorrelation && r2 >= r1 && r2 >= r3) {
            r1_best = r1;
            maximumCorrelation = r2;
            r3_best = r3;
            ir = ileft2 - 1;
            *peak = localPeak;
        }
    }
    if (maximumCorrelation > -1.0) {
        double d2r = 2 * maximumCorrelation - r1_best - r3_best;
        if (d2r != 0.0) {
            double dr = 0.5 * (r3_best - r1_best);
            maximumCorrelation += 0.5 * dr * dr / d2r;
            ir += dr / d2r;
        }
        *tout = t1 + (ir - ileft1) * me->dx;
    }
    return maximumCorrelation;
}

/* fon/Pitch_to_PointProcess.cpp Sound_Pitch_to_PointProcess_cc (mshds_extractor.py:271,321 and inside :241). */
Points *sound_pitch_to_pointprocess_cc(const Sound *sound, const Pitch *pitch) {
    Points *point = points_create(sound->xmin, sound->xmax);
    double t = pitch->xmin;
    double addedRight = -1e308;
    double globalPeak = 0.0, peak;
    for (long i = 1; i <= sound->nx; i++)
        if (fabs(Z(sound, i)) > globalPeak) globalPeak = fabs(Z(sound, i));   /* Vector_getAbsoluteExtremum, no interpolation */
    for (;;) {
        double tleft, tright;
        if (!pitch_getVoicedIntervalAfter(pitch, t, &tleft, &tright)) break;
        double tmiddle = (tleft + tright) / 2;
        double f0middle = pitch_getValueAtTime(pitch, tmiddle);
        if (isundef(f0middle)) { t = tright; continue; }   /* Praat: Melder_fatal; cannot happen for a voiced stretch */
        double tmax = sound_findExtremum(sound, tmiddle - 0.5 / f0middle, tmiddle + 0.5 / f0middle, 1, 1);
        points_add(point, tmax);
        double tsave = tmax;
        for (;;) {
            double f0 = pitch_getValueAtTime(pitch, tmax), correlation;
            if (isundef(f0)) break;
            correlation = sound_findMaximumCorrelation(sound, tmax, 1.0 / f0, tmax - 1.25 / f0, tmax - 0.8 / f0, &tmax, &peak);
            if (correlation == -1) tmax -= 1.0 / f0;
            if (tmax < tleft) {
                if (correlation > 0.7 && peak > 0.023333 * globalPeak && tmax - addedRight > 0.8 / f0) points_add(point, tmax);
                break;
            }
            if (correlation > 0.3 && (peak == 0.0 || peak > 0.01 * globalPeak)) {
                if (tmax - addedRight > 0.8 / f0) points_add(point, tmax);
            }
        }
        tmax = tsave;
        for (;;) {
            double f0 = pitch_getValueAtTime(pitch, tmax), correlation;
            if (isundef(f0)) break;
            correlation = sound_findMaximumCorrelation(sound, tmax, 1.0 / f0, tmax + 0.8 / f0, tmax + 1.25 / f0, &tmax, &peak);
            if (correlation == -1) tmax += 1.0 / f0;
            if (tmax > tright) {
                if (correlation > 0.7 && peak > 0.023333 * globalPeak) {
                    points_add(point, tmax);
                    addedRight = tmax;
                }
                break;
            }
            if (correlation > 0.3 && (peak == 0.0 || peak > 0.01 * globalPeak)) {
                points_add(point, tmax);
                addedRight = tmax;
            }
        }
        t = tright;
    }
    return point;
}
