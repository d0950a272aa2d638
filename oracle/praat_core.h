/*
 * oracle/praat_core.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU (float64, serial) restatement of the Praat 6.1.38 routines that
 * /root/reference/src/mshds_extractor.py reaches through praat-parselmouth 0.4.6
 * (pinned at /root/reference/conda-lock.yml:2683-2689; NOT vendored, NOT installed here).
 *
 * PARITY UNPINNED: neither parselmouth nor Praat exists in this image or on the GPU box, the
 * reference has no tests and no golden vectors, so this restatement is checked only
 * by analytic known-answer tests (tests/test_oracle_*.py).  Every function cites the reference call
 * site (mshds_extractor.py:line) it serves and the Praat source file whose published algorithm it
 * restates.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything under oracle/.
 *
 * Conventions: all "1-based" arrays are plain C pointers already offset by -1 by the caller
 * (y1[1..n]); Sound.z is 0-based storage and Z(s,i) is the 1-based accessor.
 */
#ifndef MSHDS_ORACLE_PRAAT_CORE_H
#define MSHDS_ORACLE_PRAAT_CORE_H

#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NUMpi 3.14159265358979323846264338327950288
#define UNDEF (NAN)
static inline int isundef(double x) { return !(x == x) || isinf(x); }
static inline int isdefined(double x) { return !isundef(x); }

/* ---- alternative readings of Praat details that cannot be checked offline ----
 * Every choice the restatement had to make without the Praat source at hand sits behind one switch; 0 is the reading all
 * committed goldens and parity tests use.  tools/appc_sensitivity.py flips them one at a time and reports which of the 25
 * columns move and by how much.  ORACLE ONLY: the CUDA path implements the default reading. */
typedef struct {
    int silence_boundary;      /* C-1  0: boundary at the frame centre x1 + (i-1) dx; 1: half a frame earlier (between frames) */
    int cut_interval;          /* ADVICE r1  0: removing a too-short interval joins it and both same-label neighbours into one
                                  (IntervalTier_cutInterval + IntervalTier_cutIntervalsOnLabelMatch); 1: the left neighbour is
                                  extended only and equal-label neighbours stay separate intervals */
    int theil_tilt_complete;   /* C-3  0: incomplete Theil (pairs i, i + n/2) in "Report spectral tilt"; 1: all pairs */
    int theil_cpps_complete;   /* C-3  same for the CPPS trend line */
    int cpps_fit_range;        /* C-4  0: qendFit = 0 <= qstartFit selects the whole quefrency domain; 1: [0.001, qmax] */
    int cpps_time_frames;      /* C-5  0: floor(0.01 / 0.002) as IEEE double division gives it (5); 1: one frame fewer (4) */
    int cpps_smooth_align;     /* C-5  0: an even box window drops its LAST tap; 1: drops its FIRST tap */
    int vuv_overlap;           /* A.12 0: overlapping V intervals stay separate rows; 1: merged; 2: a V interval starts no earlier
                                  than the previous one ends */
    int ltas_fill;             /* C-6  0: empty LTAS bands filled in increasing band order (a filled band can feed the next);
                                  1: filled from measured bands only */
    int candidate_bound;       /* C-2  0: candidate lags i < maximumLag && i < brent_ixmax; 1: i <= (inclusive bounds) */
} OrcOptions;
extern OrcOptions orc_opt;

/* ---- Sampled (fon/Sampled.cpp) ---- */
typedef struct {
    double xmin, xmax;
    long nx;
    double dx, x1;
    double *z;      /* nx samples, 0-based storage */
    int owns;
} Sound;
#define Z(s, i) ((s)->z[(i) - 1])

static inline double s_indexToX(const Sound *s, double i) { return s->x1 + (i - 1.0) * s->dx; }
static inline double s_xToIndex(const Sound *s, double x) { return (x - s->x1) / s->dx + 1.0; }
static inline long s_xToLowIndex(const Sound *s, double x) { return (long)floor(s_xToIndex(s, x)); }
static inline long s_xToHighIndex(const Sound *s, double x) { return (long)ceil(s_xToIndex(s, x)); }
static inline long iround(double x) { return (long)floor(x + 0.5); }
static inline long s_xToNearestIndex(const Sound *s, double x) { return iround(s_xToIndex(s, x)); }

Sound *sound_create(double xmin, double xmax, long nx, double dx, double x1);
Sound *sound_from_pcm16(const int16_t *pcm, long n, double fs);
Sound *sound_copy(const Sound *s);
void sound_free(Sound *s);

/* returns 0 on "shorter than window" (Praat throws) */
int shortTermAnalysis(long nx, double dx, double x1, double windowDuration, double timeStep,
                      long *numberOfFrames, double *firstTime);
long getWindowSamples(double x1, double dx, long nx, double xmin, double xmax, long *imin, long *imax);

/* ---- melder/NUM ---- */
double NUMbessel_i0_f(double x);
double NUM_interpolate_sinc(const double *y1, long n, double x, long maxDepth);
enum { PEAK_NONE = 0, PEAK_PARABOLIC = 1, PEAK_CUBIC = 2, PEAK_SINC70 = 3, PEAK_SINC700 = 4 };
double NUMimproveExtremum(const double *y1, long n, long ixmid, int interpolation, double *ixmid_real, int isMaximum);
double NUMminimize_brent(double (*f)(double, void *), double a, double b, void *closure, double tol, double *fx);
void sort_doubles(double *a, long n);
double NUMquantile(const double *a1, long n, double factor);   /* a1 sorted, 1-based */
void NUMlineFit_theil(const double *x1, const double *y1, long n, double *m, double *intercept, int complete);

/* complex FFT, n power of two, sign=-1 forward, +1 backward (unnormalised) */
void fft_pow2(double *re, double *im, long n, int sign);

/* ---- Vector queries (fon/Vector.cpp) on a contour (x1, dx, nx, y1[1..nx]) ---- */
typedef struct {
    double xmin, xmax;
    long nx;
    double dx, x1;
    double *y;      /* 0-based storage */
} Contour;
void contour_free(Contour *c);
double vector_getValueAtX(const Contour *c, double x, int interpolation /*0 nearest,1 linear,2 cubic,3 sinc70*/);
void vector_getMaximumAndX(const Contour *c, double xmin, double xmax, int interpolation, double *maximum, double *xOfMax);
void vector_getMinimumAndX(const Contour *c, double xmin, double xmax, int interpolation, double *minimum, double *xOfMin);
double contour_getQuantile(const Contour *c, double q);

/* ---- Pitch (fon/Sound_to_Pitch.cpp, fon/Pitch.cpp) ---- */
enum { AC_HANNING = 0, FCC_NORMAL = 2 };
typedef struct {
    double xmin, xmax;
    long nx;
    double dx, x1;
    double ceiling;
    int maxnCandidates;
    int *nCandidates;       /* [nx] */
    double *intensity;      /* [nx] */
    double *freq;           /* [nx*maxn], candidate 1 first */
    double *strength;       /* [nx*maxn] */
} Pitch;
void pitch_free(Pitch *p);
/* returns NULL when Praat would throw */
Pitch *sound_to_pitch_any(const Sound *me, double dt, double minimumPitch, double periodsPerWindow, int maxnCandidates,
                          int method, double silenceThreshold, double voicingThreshold, double octaveCost,
                          double octaveJumpCost, double voicedUnvoicedCost, double ceiling);
static inline int pitch_isVoiced_i(const Pitch *p, long i) {
    if (i < 1 || i > p->nx) return 0;
    double f = p->freq[(i - 1) * p->maxnCandidates];
    return f > 0.0 && f < p->ceiling;
}
double pitch_getValueAtTime(const Pitch *p, double t);     /* Hertz, linear */
double pitch_getMeanHz(const Pitch *p);
double pitch_getStdevSemitones(const Pitch *p);
int pitch_getVoicedIntervalAfter(const Pitch *p, double after, double *tleft, double *tright);

/* ---- Intensity (fon/Sound_to_Intensity.cpp) ---- */
Contour *sound_to_intensity(const Sound *me, double minimumPitch, double timeStep, int subtractMean);
double intensity_getMeanEnergy(const Contour *c);

/* ---- Silences (dwtools/Intensity_extensions.cpp) ---- */
typedef struct { double xmin, xmax; int sounding; } Interval;
typedef struct { Interval *v; long n; } Tier;
void tier_free(Tier *t);
Tier *intensity_to_silences(const Contour *me, double silenceThreshold_dB, double minSilenceDuration,
                            double minSoundingDuration);
long tier_intervalAtTime(const Tier *t, double time);   /* 1-based, 0 if outside */

/* ---- PointProcess ---- */
typedef struct { double *t; long n, cap; double xmin, xmax; } Points;
Points *points_create(double xmin, double xmax);
void points_add(Points *p, double t);
void points_free(Points *p);
Points *contour_to_points_extrema_maxima(const Contour *c, int interpolation);
Points *sound_pitch_to_pointprocess_cc(const Sound *sound, const Pitch *pitch);

/* ---- Harmonicity ---- */
/* fills *mean with mean over frames != -200 (NaN if none); returns 0 if Praat would throw */
int sound_harmonicity_cc_mean(const Sound *me, double dt, double minimumPitch, double silenceThreshold,
                              double periodsPerWindow, double *mean);

/* ---- LTAS ---- */
/* returns 0 if Praat would throw */
int pointprocess_sound_to_ltas(const Points *pulses, const Sound *sound, double maximumFrequency, double bandWidth,
                               double shortestPeriod, double longestPeriod, double maximumPeriodFactor,
                               double *ltas_dB /*[nbands]*/, long *nbands_out);
double ltas_getSlope_dB(const double *ltas0, long nx, double dx, double f1min, double f1max, double f2min, double f2max);
int ltas_fitTiltLine_robust(const double *ltas0, long nx, double dx, double fmin, double fmax, double *slope,
                            double *intercept);

/* ---- Sound manipulation ---- */
Sound *sound_upsample(const Sound *me);
Sound *sound_resample(const Sound *me, double samplingFrequency, long precision);
Sound *sound_extractPart(const Sound *me, double t1, double t2);  /* rectangular, preserveTimes=false */
void sound_preEmphasis(Sound *me, double preEmphasisFrequency);

/* ---- Formant ---- */
typedef struct {
    double xmin, xmax;
    long nx;
    double dx, x1;
    int maxnFormants;
    int *nFormants;     /* [nx] */
    double *f;          /* [nx*maxn] */
    double *bw;         /* [nx*maxn] */
} Formant;
void formant_free(Formant *f);
Formant *sound_to_formant_burg(const Sound *me, double dt, double nFormants, double maximumFrequency,
                               double halfdt_window, double preemphasisFrequency);
double formant_getValueAtTime(const Formant *me, int iformant, double t, int bandwidth);
int polynomial_roots(const double *c /* c[0..n], c[n] != 0 */, int n, double *re, double *im);
double VECburg(double *a1 /*[1..m]*/, int m, const double *x1 /*[1..n]*/, long n);

/* ---- Cepstrogram / CPPS ---- */
/* returns 0 if Praat would throw */
int sound_cpps(const Sound *me, double pitchFloor, double dt, double maximumFrequency, double preEmphasisFrequency,
               double timeAveragingWindow, double quefrencyAveragingWindow, double peakFloor, double peakCeiling,
               double qstartFit, double qendFit, double *cpps);

/* ---- Spectrogram moments ---- */
/* returns 0 if Praat would throw; mean moments over voiced frames (NaN if none) */
int sound_spectral_moments(const Sound *me, const Pitch *pitch, double effectiveAnalysisWidth, double fmax,
                           double minimumTimeStep, double minimumFreqStep, double out4[4]);

#endif
