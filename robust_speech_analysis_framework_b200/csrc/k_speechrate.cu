// k_speechrate.cu -- the de Jong & Wempe syllable-nuclei chain of _speechrate (mshds_extractor.py:11-125) on top of
// the 16 ms intensity contour and the 30-450 Hz pitch pass:
//   thresholds (:42-52), Intensity_to_TextGrid_detectSilences (:55, dwtools/Intensity_extensions.cpp), sounding table
//   (:58-73), Sound_to_PointProcess_extrema with Sinc70 refinement (:76-81), cubic value look-up (:84-88), the dip
//   filter (:91-101), voiced / sounding syllable count (:104-111) and the five ratios (:113-120).
//
// The data are tiny (62.5 contour frames per second), decisions are sequential, so one CTA owns one clip: the four
// warps share the Brent/sinc70 refinements of the intensity peaks, thread 0 runs the interval logic.
#include "internal.h"
#include "common.cuh"
#include "num.cuh"
#include "pitchq.cuh"

#define SRW 4

// Vector_getValueAtX on the contour (y1 1-based), interpolation depth 2 = cubic, 0 = nearest
__device__ double contour_value_at(const double* y1, int n, double x1, double dx, double x, int depth) {
    double leftEdge = x1 - 0.5 * dx, rightEdge = leftEdge + n * dx;
    if (x < leftEdge || x > rightEdge) return DEVNAN;
    double xi = (x - x1) / dx + 1.0;
    int midleft = (int)floor(xi), midright = midleft + 1;
    if (n < 1) return DEVNAN;
    if (xi > n) return y1[n];
    if (xi < 1) return y1[1];
    if (xi == midleft) return y1[midleft];
    int maxDepth = depth;
    if (maxDepth > midright - 1) maxDepth = midright - 1;
    if (maxDepth > n - midleft) maxDepth = n - midleft;
    if (maxDepth <= 0) return y1[(int)floor(xi + 0.5)];
    if (maxDepth == 1) return y1[midleft] + (xi - midleft) * (y1[midright] - y1[midleft]);
    double yl = y1[midleft], yr = y1[midright];
    double dyl = 0.5 * (yr - y1[midleft - 1]), dyr = 0.5 * (y1[midright + 1] - yl);
    double fil = xi - midleft, fir = midright - xi;
    return yl * fir + yr * fil - fil * fir * (0.5 * (dyr - dyl) + (fil - 0.5) * (dyl + dyr - 2 * (yr - yl)));
}

// IntervalTier_cutIntervals_minimumDuration (merge semantics as in the oracle / DESIGN.md)
__device__ void cut_short(Ivl* v, int* pn, int sounding, double minimumDuration) {
    int n = *pn, i = 0;
    while (i < n) {
        if (v[i].sounding == sounding && v[i].xmax - v[i].xmin < minimumDuration && n > 1) {
            double xmin = v[i].xmin, xmax = v[i].xmax;
            if (i == 0) {
                v[1].xmin = xmin;
                for (int j = 0; j < n - 1; j++) v[j] = v[j + 1];
                n -= 1;
            } else if (i == n - 1) {
                v[i - 1].xmax = xmax;
                n -= 1;
            } else {
                v[i - 1].xmax = v[i + 1].xmax;
                for (int j = i; j < n - 2; j++) v[j] = v[j + 2];
                n -= 2;
            }
        } else i++;
    }
    *pn = n;
}

__global__ void __launch_bounds__(SRW * 32) k_speechrate(Clips c, IntensityPass ip, const double* __restrict__ istats,
                                                          PitchPass pp, SpeechRateScratch sc, const double2* __restrict__ tw) {
    __shared__ int s_npk;
    const int clip = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = ip.nF[clip];
    const long long nx = c.off[clip + 1] - c.off[clip];
    const double xmin = 0.0, xmax = c.xmax[clip], dur_nx = (double)nx * c.dx;
    double* feat = c.feat + (size_t)clip * N_FEAT;
    // Praat throws (-> five NaNs, :124-125) when: the default to_harmonicity_cc() at :36 cannot run (FCC, 75 Hz, ppw 1:
    // needs a sound of at least 2/75 s), the intensity analysis cannot run (6.4/50 s), or the pitch analysis cannot.
    bool ok = n >= 1 && pp.nF[clip] >= 1 && !(2.0 / 75.0 > dur_nx) && !(75.0 < 1.0 / dur_nx);
    if (!ok) {
        if (threadIdx.x == 0) atomicOr(&c.status[clip], ST_SPEECHRATE);
        return;
    }
    const int base = ip.fstart[clip];
    const double* y1 = ip.out + base - 1;                 // 1-based contour
    const double x1 = ip.t1[clip], dx = ip.dt;
    const int sbase = base + 2 * clip;                    // scratch slots of this clip (capacity n + 2)
    Ivl* iv = sc.ivl + sbase;
    double* pk_t = sc.pk_t + sbase;
    double* pk_v = sc.pk_v + sbase;
    int* pk_i = sc.pk_i + sbase;

    const double silencedb = -25.0, mindip = 2.0, minpause = 0.3;
    const double min_intensity = istats[clip * 4 + 0], max_intensity = istats[clip * 4 + 1], max_99 = istats[clip * 4 + 2];
    double silencedb_1 = max_99 + silencedb;
    if (silencedb_1 < min_intensity) silencedb_1 = min_intensity;
    const double silencedb_2 = silencedb - (max_intensity - max_99);

    // ---- local maxima of the contour (ordered list), then Sinc70 refinement shared by the warps
    if (threadIdx.x == 0) {
        int k = 0;
        for (int i = 2; i <= n - 1; i++)
            if (y1[i] > y1[i - 1] && y1[i] >= y1[i + 1]) pk_i[k++] = i;
        s_npk = k;
    }
    __syncthreads();
    const int npk = s_npk;
    for (int k = warp; k < npk; k += SRW) {
        double i_real;
        (void)improve_extremum_warp(y1, n, pk_i[k], PEAK_SINC70, &i_real, true, lane, tw);
        if (lane == 0) {
            double t = x1 + (i_real - 1.0) * dx;
            pk_t[k] = t;
            pk_v[k] = contour_value_at(y1, n, x1, dx, t, 2);          // :85 "Get value at time", t, "Cubic"
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;

    // ---- Intensity_to_TextGrid_detectSilences (silencedb_2, minpause, 0.1)
    int ni = 1;
    iv[0].xmin = xmin; iv[0].xmax = xmax; iv[0].sounding = 1;
    {
        const double duration = xmax - xmin;
        const double intensityThreshold = max_intensity - fabs(silencedb_2);
        if (!(silencedb_2 < 0.0)) { atomicOr(&c.status[clip], ST_SPEECHRATE); return; }
        if (!(minpause > duration || intensityThreshold < min_intensity)) {
            int inSilence = y1[1] < intensityThreshold;
            int k = 0;
            double start = xmin;
            for (int i = 2; i <= n; i++) {
                int silent = y1[i] < intensityThreshold;
                if (silent != inSilence) {
                    double time = x1 + (i - 1) * dx;
                    iv[k].xmin = start; iv[k].xmax = time; iv[k].sounding = !inSilence;
                    k++;
                    start = time;
                    inSilence = silent;
                }
            }
            iv[k].xmin = start; iv[k].xmax = xmax; iv[k].sounding = !inSilence;
            ni = k + 1;
            cut_short(iv, &ni, 1, 0.1);
            cut_short(iv, &ni, 0, minpause);
        }
    }
    int npauses = 0;
    double Phonation_Time = 0.0, begin_speak = 0.0, end_speak = 0.0;
    for (int i = 0; i < ni; i++)
        if (iv[i].sounding) {
            if (npauses == 0) begin_speak = iv[i].xmin;
            end_speak = iv[i].xmax;
            Phonation_Time += iv[i].xmax - iv[i].xmin;
            npauses++;
        }
    if (npauses == 0) return;                                            // :63-64 five NaNs

    // ---- keep peaks above silencedb_1 (:84-88), compacting in place
    int nkept = 0;
    for (int k = 0; k < npk; k++)
        if (pk_v[k] > silencedb_1) { pk_t[nkept] = pk_t[k]; pk_v[nkept] = pk_v[k]; nkept++; }

    // ---- dip filter (:91-101) fused with the voiced / sounding test (:106-111)
    PitchView pv;
    const PitchCfg& g = pp.cfg[0];
    pv.f = pp.sel_f + pp.fstart[clip]; pv.nx = pp.nF[clip]; pv.x1 = pp.t1[clip]; pv.dx = g.dt; pv.ceiling = g.ceiling;
    pv.xmin = xmin; pv.xmax = xmax;
    int Number_Syllables = 0;
    if (nkept > 1) {
        double currenttime = pk_t[0], currentint = pk_v[0];
        for (int p = 0; p < nkept - 1; p++) {
            const double tnext = pk_t[p + 1];
            // Intensity "Get minimum" (currenttime, tnext, "None")
            double dip;
            long long imin, imax;
            if (!get_window_samples(x1, dx, n, currenttime, tnext, &imin, &imax)) {
                double yl = contour_value_at(y1, n, x1, dx, currenttime, 0), yr = contour_value_at(y1, n, x1, dx, tnext, 0);
                dip = yl < yr ? yl : yr;
            } else {
                dip = y1[imin];
                for (long long i = imin + 1; i <= imax; i++) dip = fmin(dip, y1[i]);
            }
            if (fabs(currentint - dip) > mindip) {
                const double time = pk_t[p];
                int which = -1;
                for (int i = 0; i < ni; i++)
                    if (time >= iv[i].xmin && time < iv[i].xmax) { which = i; break; }
                if (which < 0 && ni > 0 && time == iv[ni - 1].xmax) which = ni - 1;
                if (which < 0) { atomicOr(&c.status[clip], ST_SPEECHRATE); return; }   // "Get label of interval" throws
                double value = pitch_value_at(pv, time);
                if (!is_undef(value) && iv[which].sounding) Number_Syllables++;
            }
            currenttime = tnext;
            currentint = pk_v[p + 1];
        }
    }
    const double Original_Dur = end_speak - begin_speak;
    const int Number_Pauses = npauses - 1;
    const double Pause_Time = Original_Dur - Phonation_Time;
    feat[0] = Original_Dur > 0 ? (double)Number_Syllables / Original_Dur : 0.0;
    feat[1] = Phonation_Time > 0 ? (double)Number_Syllables / Phonation_Time : 0.0;
    feat[2] = Original_Dur > 0 ? Phonation_Time / Original_Dur : 0.0;
    feat[3] = Original_Dur > 0 ? (double)Number_Pauses / Original_Dur : 0.0;
    feat[4] = Number_Pauses > 0 ? Pause_Time / (double)Number_Pauses : 0.0;
}

void launch_speechrate(const Clips& c, const IntensityPass& ip, const double* istats, const PitchPass& pp,
                       const SpeechRateScratch& sc, const double2* tw, cudaStream_t s) {
    k_speechrate<<<c.n, SRW * 32, 0, s>>>(c, ip, istats, pp, sc, tw);
}
