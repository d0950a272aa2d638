/*
 * oracle/praat_pitch.c -- TEST INFRASTRUCTURE (see praat_core.h header).  PARITY UNPINNED.
 *
 * Restates Praat 6.1.38 fon/Sound_to_Pitch.cpp (Sound_to_Pitch_any, Sound_into_PitchFrame: Boersma 1993
 * autocorrelation and forward cross-correlation), fon/Pitch.cpp (Pitch_pathFinder, value/mean/sd queries,
 * Pitch_getVoicedIntervalAfter), fon/Sound_to_Harmonicity.cpp (Sound_to_Harmonicity_cc).
 * Serves mshds_extractor.py:36 (dead), :104, :143, :178, :221, :241 (inside LTAS), :270, :320, :355.
 */
#include "praat_core.h"

void pitch_free(Pitch *p) {
    if (!p) return;
    free(p->nCandidates); free(p->intensity); free(p->freq); free(p->strength);
    free(p);
}

static Pitch *pitch_create(double xmin, double xmax, long nx, double dx, double x1, double ceiling, int maxn) {
    Pitch *p = (Pitch *)calloc(1, sizeof(Pitch));
    p->xmin = xmin; p->xmax = xmax; p->nx = nx; p->dx = dx; p->x1 = x1; p->ceiling = ceiling;
    p->maxnCandidates = maxn;
    p->nCandidates = (int *)calloc((size_t)nx, sizeof(int));
    p->intensity = (double *)calloc((size_t)nx, sizeof(double));
    p->freq = (double *)calloc((size_t)nx * maxn, sizeof(double));
    p->strength = (double *)calloc((size_t)nx * maxn, sizeof(double));
    for (long i = 0; i < nx; i++) p->nCandidates[i] = 1;   /* the voiceless candidate */
    return p;
}

/* fon/Pitch.cpp Pitch_pathFinder (Viterbi). */
static void pitch_pathFinder(Pitch *me, double silenceThreshold, double voicingThreshold, double octaveCost,
                             double octaveJumpCost, double voicedUnvoicedCost, double ceiling) {
    long nx = me->nx;
    int maxn = me->maxnCandidates;
    if (nx < 1) return;
    double timeStepCorrection = 0.01 / me->dx;
    octaveJumpCost *= timeStepCorrection;
    voicedUnvoicedCost *= timeStepCorrection;
    me->ceiling = ceiling;
    double *delta = (double *)malloc(sizeof(double) * (size_t)nx * maxn);
    int *psi = (int *)calloc((size_t)nx * maxn, sizeof(int));
    for (long iframe = 0; iframe < nx; iframe++) {
        double unvoicedStrength = silenceThreshold <= 0 ? 0.0 : 2.0 - me->intensity[iframe] / (silenceThreshold / (1.0 + voicingThreshold));
        unvoicedStrength = voicingThreshold + (unvoicedStrength > 0 ? unvoicedStrength : 0);
        for (int icand = 0; icand < me->nCandidates[iframe]; icand++) {
            double f = me->freq[iframe * maxn + icand];
            int voiceless = !(f > 0.0 && f < ceiling);
            delta[iframe * maxn + icand] = voiceless ? unvoicedStrength : me->strength[iframe * maxn + icand] - octaveCost * log2(ceiling / f);
        }
    }
    for (long iframe = 1; iframe < nx; iframe++) {
        double *prevDelta = delta + (iframe - 1) * maxn, *curDelta = delta + iframe * maxn;
        int *curPsi = psi + iframe * maxn;
        for (int icand2 = 0; icand2 < me->nCandidates[iframe]; icand2++) {
            double f2 = me->freq[iframe * maxn + icand2];
            volatile double maximum = -1e30;
            int place = 0;
            for (int icand1 = 0; icand1 < me->nCandidates[iframe - 1]; icand1++) {
                double f1 = me->freq[(iframe - 1) * maxn + icand1];
                double transitionCost;
                int previousVoiceless = !(f1 > 0.0 && f1 < ceiling);
                int currentVoiceless = !(f2 > 0.0 && f2 < ceiling);
                if (currentVoiceless) {
                    transitionCost = previousVoiceless ? 0.0 : voicedUnvoicedCost;
                } else {
                    transitionCost = previousVoiceless ? voicedUnvoicedCost : octaveJumpCost * fabs(log2(f1 / f2));
                }
                volatile double value = prevDelta[icand1] - transitionCost + curDelta[icand2];
                if (value > maximum) { maximum = value; place = icand1; }
            }
            curDelta[icand2] = maximum;
            curPsi[icand2] = place;
        }
    }
    int place = 0;
    double maximum = delta[(nx - 1) * maxn];
    for (int icand = 1; icand < me->nCandidates[nx - 1]; icand++)
        if (delta[(nx - 1) * maxn + icand] > maximum) { place = icand; maximum = delta[(nx - 1) * maxn + icand]; }
    for (long iframe = nx - 1; iframe >= 0; iframe--) {
        double hf = me->freq[iframe * maxn], hs = me->strength[iframe * maxn];
        me->freq[iframe * maxn] = me->freq[iframe * maxn + place];
        me->strength[iframe * maxn] = me->strength[iframe * maxn + place];
        me->freq[iframe * maxn + place] = hf;
        me->strength[iframe * maxn + place] = hs;
        place = psi[iframe * maxn + place];
    }
    free(delta); free(psi);
}

/* fon/Sound_to_Pitch.cpp Sound_to_Pitch_any + Sound_into_PitchFrame. */
Pitch *sound_to_pitch_any(const Sound *me, double dt, double minimumPitch, double periodsPerWindow, int maxnCandidates,
                          int method, double silenceThreshold, double voicingThreshold, double octaveCost,
                          double octaveJumpCost, double voicedUnvoicedCost, double ceiling) {
    double interpolation_depth = method == AC_HANNING ? 0.5 : 1.0;
    int brent_depth = PEAK_SINC70;
    if (maxnCandidates < ceiling / minimumPitch) maxnCandidates = (int)floor(ceiling / minimumPitch);
    if (dt <= 0.0) dt = periodsPerWindow / minimumPitch / 4.0;
    double duration = me->dx * me->nx;
    if (minimumPitch < periodsPerWindow / duration) return NULL;

    long nsamp_period = (long)floor(1.0 / me->dx / minimumPitch);
    long halfnsamp_period = nsamp_period / 2 + 1;
    if (ceiling > 0.5 / me->dx) ceiling = 0.5 / me->dx;

    double dt_window = periodsPerWindow / minimumPitch;
    long nsamp_window = (long)floor(dt_window / me->dx);
    long halfnsamp_window = nsamp_window / 2 - 1;
    if (halfnsamp_window < 2) return NULL;
    nsamp_window = halfnsamp_window * 2;

    long minimumLag = (long)floor(1.0 / me->dx / ceiling);
    if (minimumLag < 2) minimumLag = 2;
    (void)minimumLag;
    long maximumLag = (long)floor(nsamp_window / periodsPerWindow) + 2;
    if (maximumLag > nsamp_window) maximumLag = nsamp_window;

    long numberOfFrames;
    double t1;
    if (!shortTermAnalysis(me->nx, me->dx, me->x1, method >= FCC_NORMAL ? 1.0 / minimumPitch + dt_window : dt_window, dt,
                           &numberOfFrames, &t1))
        return NULL;

    Pitch *thee = pitch_create(me->xmin, me->xmax, numberOfFrames, dt, t1, ceiling, maxnCandidates);

    /* global absolute peak */
    double globalPeak = 0.0;
    {
        double mean = 0.0;
        for (long i = 1; i <= me->nx; i++) mean += Z(me, i);
        mean /= me->nx;
        for (long i = 1; i <= me->nx; i++) {
            double value = fabs(Z(me, i) - mean);
            if (value > globalPeak) globalPeak = value;
        }
    }
    if (globalPeak == 0.0) return thee;

    long nsampFFT = 0, brent_ixmax;
    double *window = NULL, *windowR = NULL, *fre = NULL, *fim = NULL;
    if (method >= FCC_NORMAL) {
        brent_ixmax = (long)floor(nsamp_window * interpolation_depth);
    } else {
        nsampFFT = 1;
        while (nsampFFT < nsamp_window * (1 + interpolation_depth)) nsampFFT *= 2;
        window = (double *)calloc((size_t)nsamp_window + 1, sizeof(double));
        windowR = (double *)calloc((size_t)nsampFFT + 1, sizeof(double));
        fre = (double *)calloc((size_t)nsampFFT, sizeof(double));
        fim = (double *)calloc((size_t)nsampFFT, sizeof(double));
        for (long i = 1; i <= nsamp_window; i++) window[i] = 0.5 - 0.5 * cos(i * 2 * NUMpi / (nsamp_window + 1));
        /* normalized autocorrelation of the window */
        for (long i = 1; i <= nsamp_window; i++) fre[i - 1] = window[i];
        fft_pow2(fre, fim, nsampFFT, -1);
        for (long i = 0; i < nsampFFT; i++) { fre[i] = fre[i] * fre[i] + fim[i] * fim[i]; fim[i] = 0.0; }
        fft_pow2(fre, fim, nsampFFT, +1);
        for (long i = 1; i <= nsampFFT; i++) windowR[i] = fre[i - 1];
        for (long i = 2; i <= nsamp_window; i++) windowR[i] /= windowR[1];
        windowR[1] = 1.0;
        brent_ixmax = (long)floor(nsamp_window * interpolation_depth);
    }

    double *frame = (double *)calloc((size_t)(nsampFFT > nsamp_window ? nsampFFT : nsamp_window) + 1, sizeof(double));
    double *rbuf = (double *)calloc((size_t)(2 * nsamp_window + 1), sizeof(double));
    double *r = rbuf + nsamp_window;   /* r[-nsamp_window .. nsamp_window] */
    long *imax = (long *)calloc((size_t)maxnCandidates + 1, sizeof(long));
    int maxn = maxnCandidates;

    for (long iframe = 1; iframe <= numberOfFrames; iframe++) {
        double *cf = thee->freq + (iframe - 1) * maxn - 1;       /* 1-based candidate views */
        double *cs = thee->strength + (iframe - 1) * maxn - 1;
        double t = t1 + (iframe - 1) * dt;
        long leftSample = s_xToLowIndex(me, t), rightSample = leftSample + 1;
        long startSample, endSample;
        double localMean = 0.0;

        startSample = rightSample - nsamp_period;
        endSample = leftSample + nsamp_period;
        for (long i = startSample; i <= endSample; i++) localMean += Z(me, i);
        localMean /= 2 * nsamp_period;

        startSample = rightSample - halfnsamp_window;
        endSample = leftSample + halfnsamp_window;
        if (method < FCC_NORMAL) {
            for (long j = 1, i = startSample; j <= nsamp_window; j++) frame[j] = (Z(me, i++) - localMean) * window[j];
            for (long j = nsamp_window + 1; j <= nsampFFT; j++) frame[j] = 0.0;
        } else {
            for (long j = 1, i = startSample; j <= nsamp_window; j++) frame[j] = Z(me, i++) - localMean;
        }

        double localPeak = 0.0;
        if ((startSample = halfnsamp_window + 1 - halfnsamp_period) < 1) startSample = 1;
        if ((endSample = halfnsamp_window + halfnsamp_period) > nsamp_window) endSample = nsamp_window;
        for (long j = startSample; j <= endSample; j++) {
            double value = fabs(frame[j]);
            if (value > localPeak) localPeak = value;
        }
        thee->intensity[iframe - 1] = localPeak > globalPeak ? 1.0 : localPeak / globalPeak;

        if (method >= FCC_NORMAL) {
            double startTime = t - 0.5 * (1.0 / minimumPitch + dt_window);
            long localSpan = maximumLag + nsamp_window, localMaximumLag, offset;
            if ((startSample = s_xToLowIndex(me, startTime)) < 1) startSample = 1;
            if (localSpan > me->nx + 1 - startSample) localSpan = me->nx + 1 - startSample;
            localMaximumLag = localSpan - nsamp_window;
            offset = startSample - 1;
            const double *amp = me->z + offset - 1;    /* amp[1..] */
            double sumx2 = 0.0;
            for (long i = 1; i <= nsamp_window; i++) {
                double x = amp[i] - localMean;
                sumx2 += x * x;
            }
            double sumy2 = sumx2;
            /* stale lags (only possible at the file end) are treated as zero: per-thread buffers in Praat */
            for (long i = -nsamp_window; i <= nsamp_window; i++) r[i] = 0.0;
            r[0] = 1.0;
            for (long i = 1; i <= localMaximumLag; i++) {
                double product = 0.0;
                double y0 = amp[i] - localMean;
                double yZ = amp[i + nsamp_window] - localMean;
                sumy2 += yZ * yZ - y0 * y0;
                for (long j = 1; j <= nsamp_window; j++) {
                    double x = amp[j] - localMean;
                    double y = amp[i + j] - localMean;
                    product += x * y;
                }
                r[-i] = r[i] = product / sqrt(sumx2 * sumy2);
            }
        } else {
            for (long i = 0; i < nsampFFT; i++) { fre[i] = frame[i + 1]; fim[i] = 0.0; }
            fft_pow2(fre, fim, nsampFFT, -1);
            for (long i = 0; i < nsampFFT; i++) { fre[i] = fre[i] * fre[i] + fim[i] * fim[i]; fim[i] = 0.0; }
            fft_pow2(fre, fim, nsampFFT, +1);
            /* fre[k] = ac[k+1] */
            r[0] = 1.0;
            for (long i = 1; i <= brent_ixmax; i++) r[-i] = r[i] = fre[i] / (fre[0] * windowR[i + 1]);
        }

        thee->nCandidates[iframe - 1] = 1;
        cf[1] = 0.0; cs[1] = 0.0;
        if (localPeak == 0) continue;

        int nCand = 1;
        imax[1] = 0;
        for (long i = 2; orc_opt.candidate_bound == 1 ? (i <= maximumLag && i < brent_ixmax) : (i < maximumLag && i < brent_ixmax); i++)
            if (r[i] > 0.5 * voicingThreshold && r[i] > r[i - 1] && r[i] >= r[i + 1]) {
                int place = 0;
                double dr = 0.5 * (r[i + 1] - r[i - 1]), d2r = 2 * r[i] - r[i - 1] - r[i + 1];
                double frequencyOfMaximum = 1 / me->dx / (i + dr / d2r);
                long offset = -brent_ixmax - 1;
                double strengthOfMaximum =
                    NUM_interpolate_sinc(&r[offset], brent_ixmax - offset, 1 / me->dx / frequencyOfMaximum - offset, 30);
                if (strengthOfMaximum > 1.0) strengthOfMaximum = 1.0 / strengthOfMaximum;
                if (nCand < maxnCandidates) {
                    place = ++nCand;
                } else {
                    double weakest = 2;
                    for (int iweak = 2; iweak <= maxnCandidates; iweak++) {
                        double localStrength = cs[iweak] - octaveCost * log2(minimumPitch / cf[iweak]);
                        if (localStrength < weakest) { weakest = localStrength; place = iweak; }
                    }
                    if (strengthOfMaximum - octaveCost * log2(minimumPitch / frequencyOfMaximum) <= weakest) place = 0;
                }
                if (place) {
                    cf[place] = frequencyOfMaximum;
                    cs[place] = strengthOfMaximum;
                    imax[place] = i;
                }
            }
        thee->nCandidates[iframe - 1] = nCand;

        for (int i = 2; i <= nCand; i++) {
            double xmid, ymid;
            long offset = -brent_ixmax - 1;
            ymid = NUMimproveExtremum(&r[offset], brent_ixmax - offset, imax[i] - offset,
                                      cf[i] > 0.3 / me->dx ? PEAK_SINC700 : brent_depth, &xmid, 1);
            xmid += offset;
            cf[i] = 1.0 / me->dx / xmid;
            if (ymid > 1.0) ymid = 1.0 / ymid;
            cs[i] = ymid;
        }
    }
    free(frame); free(rbuf); free(imax); free(window); free(windowR); free(fre); free(fim);

    pitch_pathFinder(thee, silenceThreshold, voicingThreshold, octaveCost, octaveJumpCost, voicedUnvoicedCost, ceiling);
    return thee;
}

/* fon/Sampled.cpp Sampled_getValueAtX with Pitch::v_getValueAtSample (Hertz, linear):
 * pitch.get_value_at_time (mshds_extractor.py:109,364) and Pitch_getValueAtTime inside PointProcess (cc). */
double pitch_getValueAtTime(const Pitch *p, double x) {
    if (x < p->xmin || x > p->xmax) return UNDEF;
    double ireal = (x - p->x1) / p->dx + 1.0;
    long ileft = (long)floor(ireal), inear, ifar;
    double phase = ireal - ileft;
    if (phase < 0.5) { inear = ileft; ifar = ileft + 1; }
    else { ifar = ileft; inear = ileft + 1; phase = 1.0 - phase; }
    if (inear < 1 || inear > p->nx) return UNDEF;
    if (!pitch_isVoiced_i(p, inear)) return UNDEF;
    double fnear = p->freq[(inear - 1) * p->maxnCandidates];
    if (ifar < 1 || ifar > p->nx) return fnear;
    if (!pitch_isVoiced_i(p, ifar)) return fnear;
    double ffar = p->freq[(ifar - 1) * p->maxnCandidates];
    return fnear + phase * (ffar - fnear);
}

/* Pitch "Get mean 0 0 Hertz" (mshds_extractor.py:179): Sampled_getMean over the whole domain = plain mean of voiced frames. */
double pitch_getMeanHz(const Pitch *p) {
    long double sum = 0.0L;
    long n = 0;
    for (long i = 1; i <= p->nx; i++)
        if (pitch_isVoiced_i(p, i)) { sum += p->freq[(i - 1) * p->maxnCandidates]; n++; }
    return n > 0 ? (double)(sum / n) : UNDEF;
}

/* Pitch "Get standard deviation 0 0 semitones" (mshds_extractor.py:180): sd (n-1) of 12*log2(f/100). */
double pitch_getStdevSemitones(const Pitch *p) {
    long double sum = 0.0L;
    long n = 0;
    for (long i = 1; i <= p->nx; i++)
        if (pitch_isVoiced_i(p, i)) { sum += 12.0 * log2(p->freq[(i - 1) * p->maxnCandidates] / 100.0); n++; }
    if (n < 2) return UNDEF;
    double mean = (double)(sum / n);
    long double sum2 = 0.0L;
    for (long i = 1; i <= p->nx; i++)
        if (pitch_isVoiced_i(p, i)) {
            double d = 12.0 * log2(p->freq[(i - 1) * p->maxnCandidates] / 100.0) - mean;
            sum2 += d * d;
        }
    return sqrt((double)(sum2 / (n - 1)));
}

/* fon/Pitch.cpp Pitch_getVoicedIntervalAfter */
int pitch_getVoicedIntervalAfter(const Pitch *me, double after, double *tleft, double *tright) {
    long ileft = (long)ceil((after - me->x1) / me->dx + 1.0), iright;
    if (ileft > me->nx) return 0;
    if (ileft < 1) ileft = 1;
    for (; ileft <= me->nx; ileft++)
        if (pitch_isVoiced_i(me, ileft)) break;
    if (ileft > me->nx) return 0;
    for (iright = ileft; iright <= me->nx; iright++)
        if (!pitch_isVoiced_i(me, iright)) break;
    iright--;
    *tleft = me->x1 + (ileft - 1) * me->dx - 0.5 * me->dx;
    *tright = me->x1 + (iright - 1) * me->dx + 0.5 * me->dx;
    if (*tleft >= me->xmax - 0.5 * me->dx) return 0;
    if (*tleft < me->xmin) *tleft = me->xmin;
    if (*tright > me->xmax) *tright = me->xmax;
    return 1;
}

/* fon/Sound_to_Harmonicity.cpp Sound_to_Harmonicity_cc + Harmonicity "Get mean 0 0" (mshds_extractor.py:221-222). */
int sound_harmonicity_cc_mean(const Sound *me, double dt, double minimumPitch, double silenceThreshold,
                              double periodsPerWindow, double *mean) {
    Pitch *pitch = sound_to_pitch_any(me, dt, minimumPitch, periodsPerWindow, 15, FCC_NORMAL, silenceThreshold, 0.0, 0.0,
                                      0.0, 0.0, 0.5 / me->dx);
    if (!pitch) return 0;
    long double sum = 0.0L;
    long n = 0;
    for (long i = 0; i < pitch->nx; i++) {
        double f = pitch->freq[i * pitch->maxnCandidates];
        if (f == 0.0) continue;    /* -200 dB: excluded from the mean */
        double r = pitch->strength[i * pitch->maxnCandidates];
        double v = r <= 1e-15 ? -150.0 : r > 1.0 - 1e-15 ? 150.0 : 10.0 * log10(r / (1.0 - r));
        sum += v;
        n++;
    }
    *mean = n > 0 ? (double)(sum / n) : UNDEF;
    pitch_free(pitch);
    return 1;
}
