/*
 * oracle/mshds_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see praat_core.h).
 *
 * CPU restatement of /root/reference/src/mshds_extractor.py: the nine helpers (:11-376) and the per-recording
 * orchestration of extract_mshds_features (:379-459), line by line, on top of the Praat restatements in
 * praat_*.c.  One clip at a time, float64, serial per clip (clips run in parallel with OpenMP only to time the
 * CPU baseline on all host cores).
 */
#include "praat_core.h"
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

OrcOptions orc_opt = {0};

/* set one Appendix-C switch by name (see praat_core.h); returns 0 for an unknown name */
EXPORT int orc_set_option(const char *name, int value) {
#define OPT(f) if (!strcmp(name, #f)) { orc_opt.f = value; return 1; }
    OPT(silence_boundary) OPT(cut_interval) OPT(theil_tilt_complete) OPT(theil_cpps_complete) OPT(cpps_fit_range)
    OPT(cpps_time_frames) OPT(cpps_smooth_align) OPT(vuv_overlap) OPT(ltas_fill) OPT(candidate_bound)
#undef OPT
    if (!strcmp(name, "reset")) { memset(&orc_opt, 0, sizeof orc_opt); return 1; }
    return 0;
}

enum {
    ST_SPEECHRATE = 1u << 0, ST_PITCHRANGE_FALLBACK = 1u << 1, ST_PITCH = 1u << 2, ST_INTENSITY = 1u << 3,
    ST_HNR = 1u << 4, ST_LTAS = 1u << 5, ST_CPP = 1u << 6, ST_FORMANT = 1u << 7, ST_MOMENTS = 1u << 8,
    ST_FILE = 1u << 31
};

/* parselmouth defaults of Sound.to_pitch_ac / to_pitch_cc */
static Pitch *to_pitch_ac(const Sound *snd, double dt, double floor_, int maxc, double sil, double vt, double oct,
                          double jump, double vuv, double ceil_) {
    return sound_to_pitch_any(snd, dt, floor_, 3.0, maxc, AC_HANNING, sil, vt, oct, jump, vuv, ceil_);
}
static Pitch *to_pitch_ac_default(const Sound *snd, double dt, double floor_, double ceil_) {
    return to_pitch_ac(snd, dt, floor_, 15, 0.03, 0.45, 0.01, 0.35, 0.14, ceil_);
}
static Pitch *to_pitch_cc_default(const Sound *snd, double dt, double floor_, double ceil_) {
    return sound_to_pitch_any(snd, dt, floor_, 1.0, 15, FCC_NORMAL, 0.03, 0.45, 0.01, 0.35, 0.14, ceil_);
}

/* mshds_extractor.py:11-125 _speechrate */
static int speechrate(const Sound *snd, double out[5]) {
    for (int k = 0; k < 5; k++) out[k] = UNDEF;
    double silencedb = -25, mindip = 2, minpause = 0.3;
    /* :36 to_harmonicity_cc() defaults (0.01, 75, 0.1, 1.0): value unused (:37-38) but a throw aborts the group */
    double hnr;
    if (!sound_harmonicity_cc_mean(snd, 0.01, 75.0, 0.1, 1.0, &hnr)) return 0;
    if (hnr < 60) mindip = 2;

    Contour *intensity = sound_to_intensity(snd, 50, 0.016, 1);             /* :41 */
    if (!intensity) return 0;
    double min_intensity, max_intensity;
    vector_getMinimumAndX(intensity, 0, 0, PEAK_PARABOLIC, &min_intensity, NULL);   /* :42 */
    vector_getMaximumAndX(intensity, 0, 0, PEAK_PARABOLIC, &max_intensity, NULL);   /* :43 */
    double max_99_intensity = contour_getQuantile(intensity, 0.99);                 /* :47 */
    double silencedb_1 = max_99_intensity + silencedb;
    if (silencedb_1 < min_intensity) silencedb_1 = min_intensity;
    double db_adjustment = max_intensity - max_99_intensity;
    double silencedb_2 = silencedb - db_adjustment;
    int ok = 0;
    Tier *textgrid = NULL;
    Points *pp = NULL;
    Pitch *pitch = NULL;
    double *timepeaks = NULL, *intensities = NULL, *validtime = NULL;
    if (!(silencedb_2 < 0)) goto done;    /* Praat throws for a non-negative threshold */
    textgrid = intensity_to_silences(intensity, silencedb_2, minpause, 0.1);        /* :55 */
    long npauses = 0;
    double Phonation_Time = 0, begin_speak = 0, end_speak = 0;
    for (long i = 0; i < textgrid->n; i++)
        if (textgrid->v[i].sounding) {
            if (npauses == 0) begin_speak = textgrid->v[i].xmin;
            end_speak = textgrid->v[i].xmax;
            Phonation_Time += textgrid->v[i].xmax - textgrid->v[i].xmin;
            npauses++;
        }
    if (npauses == 0) { ok = 1; goto done; }   /* :63-64 returns NaNs (not an exception, same encoding) */

    pp = contour_to_points_extrema_maxima(intensity, PEAK_SINC70);                  /* :76-78 */
    long numpeaks = pp->n;
    timepeaks = (double *)malloc(sizeof(double) * (size_t)(numpeaks + 1));
    intensities = (double *)malloc(sizeof(double) * (size_t)(numpeaks + 1));
    validtime = (double *)malloc(sizeof(double) * (size_t)(numpeaks + 1));
    long nkept = 0;
    for (long i = 0; i < numpeaks; i++) {
        double value = vector_getValueAtX(intensity, pp->t[i], 2);                  /* :85 Cubic */
        if (value > silencedb_1) { intensities[nkept] = value; timepeaks[nkept] = pp->t[i]; nkept++; }
    }
    long nvalid = 0;
    if (nkept > 1) {                                                                /* :91-101 */
        double currenttime = timepeaks[0];
        double currentint = intensities[0];
        for (long p = 0; p < nkept - 1; p++) {
            long following = p + 1;
            double dip;
            vector_getMinimumAndX(intensity, currenttime, timepeaks[following], PEAK_NONE, &dip, NULL);
            if (fabs(currentint - dip) > mindip) validtime[nvalid++] = timepeaks[p];
            currenttime = timepeaks[following];
            currentint = vector_getValueAtX(intensity, timepeaks[following], 2);
        }
    }
    pitch = to_pitch_ac(snd, 0.02, 30, 4, 0.03, 0.25, 0.01, 0.35, 0.25, 450);        /* :104 */
    if (!pitch) goto done;
    long Number_Syllables = 0;
    for (long i = 0; i < nvalid; i++) {
        long whichInterval = tier_intervalAtTime(textgrid, validtime[i]);
        if (whichInterval < 1) goto done;      /* "Get label of interval" would throw */
        double value = pitch_getValueAtTime(pitch, validtime[i]);
        if (!isundef(value) && textgrid->v[whichInterval - 1].sounding) Number_Syllables++;
    }
    {
        double Original_Dur = end_speak - begin_speak;
        out[0] = Original_Dur > 0 ? Number_Syllables / Original_Dur : 0;
        out[1] = Phonation_Time > 0 ? Number_Syllables / Phonation_Time : 0;
        out[2] = Original_Dur > 0 ? Phonation_Time / Original_Dur : 0;
        long Number_Pauses = npauses - 1;
        double Pause_Time = Original_Dur - Phonation_Time;
        out[3] = Original_Dur > 0 ? Number_Pauses / Original_Dur : 0;
        out[4] = Number_Pauses > 0 ? Pause_Time / Number_Pauses : 0;
    }
    ok = 1;
done:
    contour_free(intensity); tier_free(textgrid); points_free(pp); pitch_free(pitch);
    free(timepeaks); free(intensities); free(validtime);
    return ok;
}

/* mshds_extractor.py:127-162 _pitch_values; returns 1 if the (75,500) fallback was taken */
static int pitch_values(const Sound *snd, double *pitch_floor, double *pitch_ceiling) {
    *pitch_floor = 75; *pitch_ceiling = 500;
    Pitch *pitch_wide = to_pitch_ac_default(snd, 0.005, 50, 600);                    /* :143 */
    if (!pitch_wide) return 1;
    long n = 0;
    double *v = (double *)malloc(sizeof(double) * (size_t)(pitch_wide->nx + 1));
    for (long i = 0; i < pitch_wide->nx; i++) {
        double f = pitch_wide->freq[i * pitch_wide->maxnCandidates];
        if (f != 0) v[n++] = f;
    }
    int fallback = 1;
    if (n > 0) {
        /* numpy mean / population std (pairwise summation differs from this plain loop at the 1e-16 level) */
        double mean = 0, var = 0;
        for (long i = 0; i < n; i++) mean += v[i];
        mean /= n;
        for (long i = 0; i < n; i++) var += (v[i] - mean) * (v[i] - mean);
        double sd = sqrt(var / n);
        double s2 = 0; long m = 0;
        for (long i = 0; i < n; i++) {
            double z = (v[i] - mean) / sd;
            if (fabs(z) <= 2) { s2 += v[i]; m++; }
        }
        if (m > 0) {
            double mean_pitch = s2 / m;
            if (mean_pitch < 170) { *pitch_floor = 60; *pitch_ceiling = 250; }
            else { *pitch_floor = 100; *pitch_ceiling = 500; }
            fallback = 0;
        }
    }
    free(v);
    pitch_free(pitch_wide);
    return fallback;
}

static double sd_ddof1(const double *v, long n, double *mean_out) {
    if (n < 1) { *mean_out = UNDEF; return UNDEF; }
    double mean = 0;
    for (long i = 0; i < n; i++) mean += v[i];
    mean /= n;
    *mean_out = mean;
    if (n < 2) return UNDEF;
    double s = 0;
    for (long i = 0; i < n; i++) s += (v[i] - mean) * (v[i] - mean);
    return sqrt(s / (n - 1));
}

/* mshds_extractor.py:253-301 _extract_CPP */
static double extract_CPP(const Sound *snd, double floor_, double ceiling) {
    Pitch *pitch = to_pitch_ac(snd, 0.005, floor_, 15, 0.03, 0.3, 0.01, 0.35, 0.14, ceiling);     /* :270 */
    if (!pitch) return UNDEF;
    Points *pulses = sound_pitch_to_pointprocess_cc(snd, pitch);                                 /* :271 */
    pitch_free(pitch);
    /* :272 PointProcess_to_TextGrid_vuv (0.02, 0.1): V intervals only */
    double maxT = 0.02, meanT = 0.1, halfMeanT = 0.5 * meanT;
    double sum = 0; long cnt = 0; int fail = 0;
    long ipointright;
    double prevEnd = snd->xmin;
    for (long ipointleft = 1; ipointleft <= pulses->n; ipointleft = ipointright + 1) {
        for (ipointright = ipointleft + 1; ipointright <= pulses->n; ipointright++)
            if (pulses->t[ipointright - 1] - pulses->t[ipointright - 2] > maxT) break;
        ipointright--;
        double beginVoiced = pulses->t[ipointleft - 1] - halfMeanT;
        if (beginVoiced < snd->xmin) beginVoiced = snd->xmin;
        double endVoiced = pulses->t[ipointright - 1] + halfMeanT;
        if (endVoiced > snd->xmax) endVoiced = snd->xmax;
        if (orc_opt.vuv_overlap == 1) {
            /* alternative reading: runs whose extended intervals overlap form ONE V interval */
            while (ipointright < pulses->n) {
                double nextBegin = pulses->t[ipointright] - halfMeanT;
                if (nextBegin > endVoiced) break;
                long j;
                for (j = ipointright + 2; j <= pulses->n; j++)
                    if (pulses->t[j - 1] - pulses->t[j - 2] > maxT) break;
                ipointright = j - 1;
                endVoiced = pulses->t[ipointright - 1] + halfMeanT;
                if (endVoiced > snd->xmax) endVoiced = snd->xmax;
            }
        } else if (orc_opt.vuv_overlap == 2) {
            /* alternative reading: a V interval cannot start before the previous one ended */
            if (beginVoiced < prevEnd) beginVoiced = prevEnd;
        }
        prevEnd = endVoiced;
        /* :273 "Down to Table" renders times with 6 decimals; :280-281 float() of the strings */
        char buf[64];
        snprintf(buf, sizeof buf, "%.6f", beginVoiced); double tmin = strtod(buf, NULL);
        snprintf(buf, sizeof buf, "%.6f", endVoiced); double tmax = strtod(buf, NULL);
        if (tmin >= tmax) continue;                                                              /* :284 */
        Sound *seg = sound_extractPart(snd, tmin, tmax);                                         /* :286 */
        if (!seg) { fail = 1; break; }     /* exception outside the inner try -> whole group NaN (:300) */
        double cpp;
        int ok = sound_cpps(seg, 60, 0.002, 5000, 50, 0.01, 0.001, 60, 330, 0.001, 0, &cpp);     /* :289,291 */
        sound_free(seg);
        if (!ok) { fail = 1; break; }
        if (!isundef(cpp) && cpp > 4) { sum += cpp; cnt++; }                                     /* :293 */
    }
    points_free(pulses);
    if (fail) return UNDEF;
    return cnt > 0 ? sum / cnt : UNDEF;
}

/* mshds_extractor.py:303-338 _measureFormants */
static void measureFormants(const Sound *snd, double floor_, double ceiling, double out[8]) {
    for (int k = 0; k < 8; k++) out[k] = UNDEF;
    Formant *formants = sound_to_formant_burg(snd, 0.005, 5, 5000, 0.025, 50);       /* :319 */
    if (!formants) return;
    Pitch *pitch = to_pitch_cc_default(snd, 0.005, floor_, ceiling);                 /* :320 */
    if (!pitch) { formant_free(formants); return; }
    Points *pulses = sound_pitch_to_pointprocess_cc(snd, pitch);                     /* :321 */
    double *lists[4];
    long cnt[4] = {0, 0, 0, 0};
    for (int k = 0; k < 4; k++) lists[k] = (double *)malloc(sizeof(double) * (size_t)(pulses->n + 1));
    for (long p = 0; p < pulses->n; p++) {
        double t = pulses->t[p], v;
        if (!isundef(v = formant_getValueAtTime(formants, 1, t, 0))) lists[0][cnt[0]++] = v;
        if (!isundef(v = formant_getValueAtTime(formants, 1, t, 1))) lists[1][cnt[1]++] = v;
        if (!isundef(v = formant_getValueAtTime(formants, 2, t, 0))) lists[2][cnt[2]++] = v;
        if (!isundef(v = formant_getValueAtTime(formants, 2, t, 1))) lists[3][cnt[3]++] = v;
    }
    for (int k = 0; k < 4; k++) {
        double mean, sd = sd_ddof1(lists[k], cnt[k], &mean);
        out[2 * k] = mean; out[2 * k + 1] = sd;
        free(lists[k]);
    }
    points_free(pulses); pitch_free(pitch); formant_free(formants);
}

/* mshds_extractor.py:379-459 body of the per-file loop, after load / mono / resample(16000, 50) */
static uint32_t extract_clip(const Sound *snd_in, double out[25]) {
    uint32_t st = 0;
    for (int k = 0; k < 25; k++) out[k] = UNDEF;
    Sound *res = NULL;
    const Sound *snd = snd_in;
    if (1.0 / snd_in->dx != 16000.0) {                                               /* :418-419 */
        res = sound_resample(snd_in, 16000, 50);
        if (!res) return ST_FILE;
        snd = res;
    }
    if (!speechrate(snd, out + 0)) st |= ST_SPEECHRATE;                               /* :426 */
    double floor_, ceiling;
    if (pitch_values(snd, &floor_, &ceiling)) st |= ST_PITCHRANGE_FALLBACK;          /* :428 */

    Pitch *pitch = to_pitch_ac_default(snd, 0.005, floor_, ceiling);                 /* :178 (and :355) */
    if (pitch) {
        out[5] = pitch_getMeanHz(pitch);                                             /* :179 */
        out[6] = pitch_getStdevSemitones(pitch);                                     /* :180 */
    } else st |= ST_PITCH;

    Contour *inten = sound_to_intensity(snd, floor_, 0.005, 1);                       /* :198 */
    if (inten) {
        double mn, mx;
        out[7] = intensity_getMeanEnergy(inten);                                     /* :199 */
        vector_getMinimumAndX(inten, 0, 0, PEAK_PARABOLIC, &mn, NULL);               /* :200 */
        vector_getMaximumAndX(inten, 0, 0, PEAK_PARABOLIC, &mx, NULL);               /* :201 */
        out[8] = mn != 0 ? mx / mn : UNDEF;                                          /* :202 */
        contour_free(inten);
    } else st |= ST_INTENSITY;

    if (!sound_harmonicity_cc_mean(snd, 0.005, floor_, 0.1, 4.5, &out[9])) st |= ST_HNR;   /* :221-222 */

    {   /* :241-248 _extract_Slope_Tilt */
        Pitch *p2 = to_pitch_ac_default(snd, 0.0, floor_, ceiling);     /* Sound_to_PointProcess_periodic_cc */
        int ok = 0;
        if (p2) {
            Points *pulses = sound_pitch_to_pointprocess_cc(snd, p2);
            double ltas[64]; long nb;
            if (pointprocess_sound_to_ltas(pulses, snd, 5000, 100, 0.0001, 0.02, 1.3, ltas, &nb)) {
                double slope, icpt;
                out[10] = ltas_getSlope_dB(ltas, nb, 100, 50, 1000, 1000, 4000);
                if (ltas_fitTiltLine_robust(ltas, nb, 100, 100, 5000, &slope, &icpt)) { out[11] = slope; ok = 1; }
                else out[10] = UNDEF;
            }
            points_free(pulses);
            pitch_free(p2);
        }
        if (!ok) st |= ST_LTAS;
    }

    out[12] = extract_CPP(snd, floor_, ceiling);                                     /* :434 */
    if (isundef(out[12])) st |= ST_CPP;

    measureFormants(snd, floor_, ceiling, out + 13);                                 /* :441 */
    if (isundef(out[13])) st |= ST_FORMANT;

    if (pitch) {                                                                     /* :446 (pitch identical to :178) */
        if (!sound_spectral_moments(snd, pitch, 0.025, 5000, 0.005, 20, out + 21)) st |= ST_MOMENTS;
        pitch_free(pitch);
    } else st |= ST_MOMENTS;
    sound_free(res);
    return st;
}

/* ------------------------------------------------------------------ exported entry points (ctypes) */

EXPORT int orc_extract(const int16_t *pcm, const int64_t *offsets, int n_clips, double fs, double *out, uint32_t *status,
                       int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int c = 0; c < n_clips; c++) {
        long n = (long)(offsets[c + 1] - offsets[c]);
        if (n <= 0) {
            for (int k = 0; k < 25; k++) out[c * 25 + k] = UNDEF;
            if (status) status[c] = ST_FILE;
            continue;
        }
        Sound *snd = sound_from_pcm16(pcm + offsets[c], n, fs);
        uint32_t st = extract_clip(snd, out + (size_t)c * 25);
        if (status) status[c] = st;
        sound_free(snd);
    }
    return 0;
}

static Sound wrap(const double *x, long n, double fs) {
    Sound s;
    s.xmin = 0; s.dx = 1.0 / fs; s.xmax = n * s.dx; s.nx = n; s.x1 = 0.5 * s.dx; s.z = (double *)x; s.owns = 0;
    return s;
}

EXPORT int orc_extract_f64(const double *x, long n, double fs, double *out25, uint32_t *status) {
    Sound s = wrap(x, n, fs);
    uint32_t st = extract_clip(&s, out25);
    if (status) *status = st;
    return 0;
}

EXPORT int orc_frame_grid(long nx, double fs, double windowDuration, double timeStep, long *nframes, double *t1) {
    double dx = 1.0 / fs;
    return shortTermAnalysis(nx, dx, 0.5 * dx, windowDuration, timeStep, nframes, t1);
}

EXPORT long orc_intensity(const double *x, long n, double fs, double minPitch, double dt, int subtractMean, double *out,
                          long cap, double *x1) {
    Sound s = wrap(x, n, fs);
    Contour *c = sound_to_intensity(&s, minPitch, dt, subtractMean);
    if (!c) return -1;
    long nf = c->nx;
    for (long i = 0; i < nf && i < cap; i++) out[i] = c->y[i];
    if (x1) *x1 = c->x1;
    contour_free(c);
    return nf;
}

/* query order: [min_parabolic, max_parabolic, q99, mean_energy] of an intensity contour */
EXPORT int orc_intensity_stats(const double *x, long n, double fs, double minPitch, double dt, double out4[4]) {
    Sound s = wrap(x, n, fs);
    Contour *c = sound_to_intensity(&s, minPitch, dt, 1);
    if (!c) return 0;
    vector_getMinimumAndX(c, 0, 0, PEAK_PARABOLIC, &out4[0], NULL);
    vector_getMaximumAndX(c, 0, 0, PEAK_PARABOLIC, &out4[1], NULL);
    out4[2] = contour_getQuantile(c, 0.99);
    out4[3] = intensity_getMeanEnergy(c);
    contour_free(c);
    return 1;
}

/* Selected path (candidate 1 after Viterbi) of a pitch analysis; method 0 = AC, 2 = FCC. */
EXPORT long orc_pitch(const double *x, long n, double fs, int method, double dt, double floor_, double ppw, int maxc,
                      double sil, double vt, double oct, double jump, double vuv, double ceil_, double *freq,
                      double *strength, int *ncand, long cap, double *x1, double *dt_out) {
    Sound s = wrap(x, n, fs);
    Pitch *p = sound_to_pitch_any(&s, dt, floor_, ppw, maxc, method, sil, vt, oct, jump, vuv, ceil_);
    if (!p) return -1;
    long nf = p->nx;
    for (long i = 0; i < nf && i < cap; i++) {
        if (freq) freq[i] = p->freq[i * p->maxnCandidates];
        if (strength) strength[i] = p->strength[i * p->maxnCandidates];
        if (ncand) ncand[i] = p->nCandidates[i];
    }
    if (x1) *x1 = p->x1;
    if (dt_out) *dt_out = p->dx;
    pitch_free(p);
    return nf;
}

EXPORT long orc_pulses(const double *x, long n, double fs, int method, double dt, double floor_, double ppw, double vt,
                       double ceil_, double *t, long cap) {
    Sound s = wrap(x, n, fs);
    Pitch *p = sound_to_pitch_any(&s, dt, floor_, ppw, 15, method, 0.03, vt, 0.01, 0.35, 0.14, ceil_);
    if (!p) return -1;
    Points *pp = sound_pitch_to_pointprocess_cc(&s, p);
    long np = pp->n;
    for (long i = 0; i < np && i < cap; i++) t[i] = pp->t[i];
    points_free(pp); pitch_free(p);
    return np;
}

EXPORT int orc_pitch_values(const double *x, long n, double fs, double *floor_, double *ceil_) {
    Sound s = wrap(x, n, fs);
    return pitch_values(&s, floor_, ceil_);
}

EXPORT int orc_speechrate(const double *x, long n, double fs, double out5[5]) {
    Sound s = wrap(x, n, fs);
    return speechrate(&s, out5);
}

EXPORT int orc_hnr(const double *x, long n, double fs, double dt, double floor_, double sil, double ppw, double *mean) {
    Sound s = wrap(x, n, fs);
    return sound_harmonicity_cc_mean(&s, dt, floor_, sil, ppw, mean);
}

EXPORT int orc_ltas(const double *x, long n, double fs, double floor_, double ceil_, double *ltas50, double out2[2]) {
    Sound s = wrap(x, n, fs);
    Pitch *p2 = to_pitch_ac_default(&s, 0.0, floor_, ceil_);
    if (!p2) return 0;
    Points *pulses = sound_pitch_to_pointprocess_cc(&s, p2);
    long nb;
    double tmp[64];
    int ok = pointprocess_sound_to_ltas(pulses, &s, 5000, 100, 0.0001, 0.02, 1.3, tmp, &nb);
    if (ok) {
        double icpt;
        if (ltas50) memcpy(ltas50, tmp, sizeof(double) * (size_t)nb);
        out2[0] = ltas_getSlope_dB(tmp, nb, 100, 50, 1000, 1000, 4000);
        ok = ltas_fitTiltLine_robust(tmp, nb, 100, 100, 5000, &out2[1], &icpt);
    }
    points_free(pulses); pitch_free(p2);
    return ok;
}

EXPORT long orc_resample(const double *x, long n, double fs, double newfs, long precision, double *out, long cap,
                         double *x1) {
    Sound s = wrap(x, n, fs);
    Sound *r = sound_resample(&s, newfs, precision);
    if (!r) return -1;
    long m = r->nx;
    for (long i = 0; i < m && i < cap; i++) out[i] = r->z[i];
    if (x1) *x1 = r->x1;
    sound_free(r);
    return m;
}

EXPORT long orc_formants(const double *x, long n, double fs, double *f, double *bw, int *nf, long cap, double *x1) {
    Sound s = wrap(x, n, fs);
    Formant *fm = sound_to_formant_burg(&s, 0.005, 5, 5000, 0.025, 50);
    if (!fm) return -1;
    long nfr = fm->nx;
    for (long i = 0; i < nfr && i < cap; i++) {
        nf[i] = fm->nFormants[i];
        for (int k = 0; k < 5; k++) {
            f[i * 5 + k] = k < fm->nFormants[i] ? fm->f[i * fm->maxnFormants + k] : UNDEF;
            bw[i * 5 + k] = k < fm->nFormants[i] ? fm->bw[i * fm->maxnFormants + k] : UNDEF;
        }
    }
    if (x1) *x1 = fm->x1;
    formant_free(fm);
    return nfr;
}

EXPORT int orc_formant_stats(const double *x, long n, double fs, double floor_, double ceil_, double out8[8]) {
    Sound s = wrap(x, n, fs);
    measureFormants(&s, floor_, ceil_, out8);
    return 1;
}

EXPORT double orc_cpp(const double *x, long n, double fs, double floor_, double ceil_) {
    Sound s = wrap(x, n, fs);
    return extract_CPP(&s, floor_, ceil_);
}

EXPORT int orc_cpps_segment(const double *x, long n, double fs, double *cpps) {
    Sound s = wrap(x, n, fs);
    return sound_cpps(&s, 60, 0.002, 5000, 50, 0.01, 0.001, 60, 330, 0.001, 0, cpps);
}

EXPORT int orc_moments(const double *x, long n, double fs, double floor_, double ceil_, double out4[4]) {
    Sound s = wrap(x, n, fs);
    Pitch *p = to_pitch_ac_default(&s, 0.005, floor_, ceil_);
    if (!p) return 0;
    int ok = sound_spectral_moments(&s, p, 0.025, 5000, 0.005, 20, out4);
    pitch_free(p);
    return ok;
}

/* numerics for unit tests */
EXPORT double orc_interpolate_sinc(const double *y, long n, double x, long depth) { return NUM_interpolate_sinc(y - 1, n, x, depth); }
EXPORT double orc_improve_extremum(const double *y, long n, long ixmid, int interpolation, int isMaximum, double *ixreal) {
    return NUMimproveExtremum(y - 1, n, ixmid, interpolation, ixreal, isMaximum);
}
EXPORT double orc_bessel_i0(double x) { return NUMbessel_i0_f(x); }
EXPORT double orc_quantile(const double *sorted, long n, double q) { return NUMquantile(sorted - 1, n, q); }
EXPORT void orc_theil(const double *x, const double *y, long n, int complete, double *m, double *b) {
    NUMlineFit_theil(x - 1, y - 1, n, m, b, complete);
}
EXPORT double orc_burg(const double *x, long n, int m, double *a) { return VECburg(a - 1, m, x - 1, n); }
EXPORT int orc_roots(const double *c, int n, double *re, double *im) { return polynomial_roots(c, n, re, im); }
EXPORT void orc_fft(double *re, double *im, long n, int sign) { fft_pow2(re, im, n, sign); }
EXPORT long orc_silences(const double *contour, long nx, double dx, double x1, double xmin, double xmax, double thr,
                         double minSil, double minSnd, double *bounds /*[cap*2]*/, int *sounding, long cap) {
    Contour c;
    c.xmin = xmin; c.xmax = xmax; c.nx = nx; c.dx = dx; c.x1 = x1; c.y = (double *)contour;
    Tier *t = intensity_to_silences(&c, thr, minSil, minSnd);
    long n = t->n;
    for (long i = 0; i < n && i < cap; i++) { bounds[2 * i] = t->v[i].xmin; bounds[2 * i + 1] = t->v[i].xmax; sounding[i] = t->v[i].sounding; }
    tier_free(t);
    return n;
}
