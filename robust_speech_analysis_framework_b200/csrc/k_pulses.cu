// k_pulses.cu -- glottal pulse detection: Sound_Pitch_to_PointProcess_cc (fon/Pitch_to_PointProcess.cpp) with
// Pitch_getVoicedIntervalAfter, Sound_findExtremum and Sound_findMaximumCorrelation.
//
// Serves "To PointProcess (cc)" at mshds_extractor.py:271 (CPP), :321 (formants) and inside "To Ltas (pitch-corrected)"
// at :241.  The reference walks every voiced stretch pulse by pulse on one thread; here every voiced stretch of every
// clip is an independent work item owned by one warp: the 32 lanes evaluate the ~0.45*T candidate offsets of
// Sound_findMaximumCorrelation in parallel (each lane runs the reference's serial sum for its offsets, so the
// correlations are bit-identical to a float64 CPU loop), and the only cross-stretch dependency (the `addedRight`
// guard) is applied afterwards by a per-clip pass that also restores time order.
#include "internal.h"
#include "common.cuh"
#include "pitchq.cuh"

#define PW 4                    // warps per CTA
#define SPAN_MAX 768           // doubles of staged samples per warp
#define OFF_MAX 288             // candidate offsets per search

// ------------------------------------------------------------------------------------------------ voiced stretches
// Pitch_getVoicedIntervalAfter applied repeatedly = maximal runs of voiced frames.  One warp per clip, in order.
__global__ void __launch_bounds__(128) k_stretch_list(Clips c, PitchPass p, PulseSet ps) {
    const int lane = threadIdx.x & 31;
    const int clip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (clip >= c.n) return;
    const int nF = p.nF[clip], f0 = p.fstart[clip];
    const double ceiling = p.cfg[c.cls[clip]].ceiling;
    int count = 0, start = 0;
    bool inrun = false;
    unsigned prev = 0;
    for (int base = 0; base < nF; base += 32) {
        int i = base + lane;
        bool v = false;
        if (i < nF) { double f = p.sel_f[f0 + i]; v = f > 0.0 && f < ceiling; }
        unsigned m = __ballot_sync(FULL_MASK, v);
        if (lane == 0) {
            unsigned x = m ^ ((m << 1) | prev);       // bit b set: frame base+b differs from its predecessor
            while (x) {
                int b = __ffs(x) - 1;
                x &= x - 1;
                if (!inrun) { start = base + b + 1; inrun = true; }                 // 1-based first voiced frame
                else { ps.st_ileft[f0 + count] = start; ps.st_iright[f0 + count] = base + b; count++; inrun = false; }
            }
            prev = m >> 31;
        }
    }
    if (lane == 0) {
        if (inrun) { ps.st_ileft[f0 + count] = start; ps.st_iright[f0 + count] = nF; count++; }
        ps.st_count[clip] = count;
        ps.valid[clip] = nF >= 1;
    }
}

// ------------------------------------------------------------------------------------------------ helpers
struct WarpSound {
    SPtr pcm;               // clip samples, pcm[i-1] = sample i
    long long nx;
    double x1, dx;
    double* stage;          // per-warp shared staging buffer [SPAN_MAX]
    double* rbuf;           // [OFF_MAX + 4]
    double* pbuf;           // [OFF_MAX + 4]
};

// Sound_findExtremum (includeMaxima = includeMinima = true) -> time of the absolute extremum, parabolically refined
__device__ double find_extremum_warp(const WarpSound& S, double tmin, double tmax, int lane) {
    long long imin = x_to_low(S.x1, S.dx, tmin), imax = x_to_high(S.x1, S.dx, tmax);
    if (imin < 1) imin = 1;
    if (imax > S.nx) imax = S.nx;
    long long n = imax - imin + 1;
    double iextremum;
    const SPtr ch = S.pcm + ((imin - 1) - 1);             // ch[i], i = 1..n  -> sample imin-1+i
    if (n < 3) {
        if (n <= 0) iextremum = 0.0;
        else if (n == 1) iextremum = 1.0;
        else {
            double a = fabs(samp(ch, 1)), b = fabs(samp(ch, 2));
            iextremum = a > b ? 1.0 : a < b ? 2.0 : 1.5;
        }
    } else {
        double mn = CUDART_INF, mx = -CUDART_INF;
        long long imn = 0x7fffffffffffLL, imx = 0x7fffffffffffLL;
        for (long long i = 1 + lane; i <= n; i += 32) {
            double v = samp(ch, i);
            if (v < mn) { mn = v; imn = i; }
            if (v > mx) { mx = v; imx = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double omn = __shfl_xor_sync(FULL_MASK, mn, o), omx = __shfl_xor_sync(FULL_MASK, mx, o);
            long long oimn = __shfl_xor_sync(FULL_MASK, imn, o), oimx = __shfl_xor_sync(FULL_MASK, imx, o);
            if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
            if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
        }
        if (mn == mx) iextremum = 0.5 * ((double)n + 1.0);
        else {
            long long iextr = fabs(mn) > fabs(mx) ? imn : imx;
            if (iextr == 1) iextremum = 1.0;
            else if (iextr == n) iextremum = (double)n;
            else {
                double vm = samp(ch, iextr), vl = samp(ch, iextr - 1), vr = samp(ch, iextr + 1);
                iextremum = (double)iextr + 0.5 * (vr - vl) / (2 * vm - vl - vr);
            }
        }
    }
    if (iextremum != 0.0) return S.x1 + ((double)(imin - 1) + iextremum - 1.0) * S.dx;
    return (tmin + tmax) / 2;
}

// Sound_findMaximumCorrelation; returns the (parabolically improved) maximum correlation or -1; *tout, *peak as Praat.
__device__ double find_max_corr_warp(const WarpSound& S, double t1, double windowLength, double tmin2, double tmax2,
                                     double* tout, double* peak, int lane) {
    const double halfWindowLength = 0.5 * windowLength;
    const long long ileft1 = x_to_nearest(S.x1, S.dx, t1 - halfWindowLength);
    const long long iright1 = x_to_nearest(S.x1, S.dx, t1 + halfWindowLength);
    const long long ileft2min = x_to_low(S.x1, S.dx, tmin2 - halfWindowLength);
    const long long ileft2max = x_to_high(S.x1, S.dx, tmax2 - halfWindowLength);
    *peak = 0.0;
    long long m_ll = ileft2max - ileft2min + 1;
    if (m_ll < 1) return -1.0;
    int m = (int)(m_ll > OFF_MAX ? OFF_MAX : m_ll);       // (never truncated for floor >= 50 Hz at 16 kHz)
    const long long len = iright1 - ileft1 + 1;
    // stage the union span [lo, hi] as float64
    long long lo = ileft1 < ileft2min ? ileft1 : ileft2min;
    long long hi1 = iright1, hi2 = ileft2min + (m - 1) + (len - 1);
    long long hi = hi1 > hi2 ? hi1 : hi2;
    const bool staged = hi - lo + 1 <= SPAN_MAX;
    __syncwarp();
    if (staged)
        for (long long i = lo + lane; i <= hi; i += 32) S.stage[i - lo] = (i >= 1 && i <= S.nx) ? samp(S.pcm, i - 1) : 0.0;
    __syncwarp();
    // norm1 (the energy of the fixed window) is the same serial sum for every offset whose range is not clipped by the
    // ends of the sound: it is accumulated once per lane (same order, same bits) and reused
    double n1_full = 0.0;
    bool have_n1 = false;
    for (int j = lane; j < m; j += 32) {
        const long long ileft2 = ileft2min + j;
        double norm1 = 0.0, norm2 = 0.0, product = 0.0, localPeak = 0.0;
        if (staged) {
            long long kA = 0, kB = len - 1;
            if (1 - ileft1 > kA) kA = 1 - ileft1;
            if (1 - ileft2 > kA) kA = 1 - ileft2;
            if (S.nx - ileft1 < kB) kB = S.nx - ileft1;
            if (S.nx - ileft2 < kB) kB = S.nx - ileft2;
            const double* p1 = S.stage + (ileft1 - lo);
            const double* p2 = S.stage + (ileft2 - lo);
            const bool full = kA == 0 && kB == len - 1;
            if (full && have_n1) {
                norm1 = n1_full;
                for (long long k = 0; k < len; k++) {
                    const double amp1 = p1[k], amp2 = p2[k];
                    norm2 += amp2 * amp2;
                    product += amp1 * amp2;
                    localPeak = fmax(localPeak, fabs(amp2));
                }
            } else {
                for (long long k = kA; k <= kB; k++) {
                    const double amp1 = p1[k], amp2 = p2[k];
                    norm1 += amp1 * amp1;
                    norm2 += amp2 * amp2;
                    product += amp1 * amp2;
                    localPeak = fmax(localPeak, fabs(amp2));
                }
                if (full) { n1_full = norm1; have_n1 = true; }
            }
        } else {
            for (long long k = 0; k < len; k++) {
                long long i1 = ileft1 + k, i2 = ileft2 + k;
                if (i1 < 1 || i1 > S.nx || i2 < 1 || i2 > S.nx) continue;
                double amp1 = samp(S.pcm, i1 - 1), amp2 = samp(S.pcm, i2 - 1);
                norm1 += amp1 * amp1;
                norm2 += amp2 * amp2;
                product += amp1 * amp2;
                if (fabs(amp2) > localPeak) localPeak = fabs(amp2);
            }
        }
        S.rbuf[j + 2] = product != 0.0 ? product / (sqrt(norm1 * norm2)) : 0.0;
        S.pbuf[j + 2] = localPeak;
    }
    if (lane == 0) { S.rbuf[0] = 0.0; S.rbuf[1] = 0.0; }   // r1 = r2 = 0 before the first offset
    __syncwarp();
    // e[j] = rbuf[j+2]; candidates j = -1 .. m-2 are tested as r2 with r1 = e[j-1], r3 = e[j+1]
    double best = -1.0;
    int bestj = 0x7fffffff;
    for (int j = -1 + lane; j <= m - 2; j += 32) {
        double r2 = S.rbuf[j + 2], r1 = S.rbuf[j + 1], r3 = S.rbuf[j + 3];
        if (r2 > best && r2 >= r1 && r2 >= r3) { best = r2; bestj = j; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(FULL_MASK, best, o);
        int oj = __shfl_xor_sync(FULL_MASK, bestj, o);
        if (ob > best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
    }
    double maximumCorrelation = -1.0;
    if (bestj != 0x7fffffff && best > -1.0) {
        maximumCorrelation = best;
        double r1_best = S.rbuf[bestj + 1], r3_best = S.rbuf[bestj + 3];
        double ir = (double)(ileft2min + bestj);
        *peak = S.pbuf[bestj + 3];                          // Praat records the peak of the window AFTER the best one
        double d2r = 2 * maximumCorrelation - r1_best - r3_best;
        if (d2r != 0.0) {
            double dr = 0.5 * (r3_best - r1_best);
            maximumCorrelation += 0.5 * dr * dr / d2r;
            ir += dr / d2r;
        }
        *tout = t1 + (ir - (double)ileft1) * S.dx;
    }
    __syncwarp();
    return maximumCorrelation;
}

// ------------------------------------------------------------------------------------------------ per-stretch walk
// Work item = (voiced stretch, direction): the left and the right walk of a stretch start from the same first pulse and do
// not depend on each other (the only coupling, the `addedRight` guard, is applied afterwards per clip), so they run on two
// warps -- half the latency of the longest chain, which is what a batch of few long recordings waits for.  Raw points of
// a stretch: left walk at raw[2 * region .. + cap), first pulse + right walk at raw[2 * region + cap .. + cap).
__global__ void __launch_bounds__(PW * 32) k_pulses_stretch(Clips c, PitchPass p, PulseSet ps) {
    __shared__ double s_stage[PW][SPAN_MAX];
    __shared__ double s_r[PW][OFF_MAX + 4];
    __shared__ double s_p[PW][OFF_MAX + 4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * PW + wib, nw = gridDim.x * PW;
    const int total = 2 * ps.st_start[c.n];
    for (int w2 = gw; w2 < total; w2 += nw) {
        const int w = w2 >> 1;
        const bool go_right = (w2 & 1) != 0;
        const int clip = find_segment(ps.st_start, c.n, w);
        const int k = w - ps.st_start[clip];
        const int slot = p.fstart[clip] + k;
        const int ileft = ps.st_ileft[slot], iright = ps.st_iright[slot];
        const PitchCfg& g = p.cfg[c.cls[clip]];
        PitchView pv;
        pv.f = p.sel_f + p.fstart[clip]; pv.nx = p.nF[clip]; pv.x1 = p.t1[clip]; pv.dx = g.dt; pv.ceiling = g.ceiling;
        pv.xmin = 0.0;
        WarpSound S;
        S.nx = c.off[clip + 1] - c.off[clip];
        S.pcm = c.pcm + c.off[clip];
        S.dx = c.dx; S.x1 = c.x1[clip];
        S.stage = s_stage[wib]; S.rbuf = s_r[wib]; S.pbuf = s_p[wib];
        pv.xmax = c.xmax[clip];
        const double globalPeak = c.apeak[clip];

        double tleft = pv.x1 + (double)(ileft - 1) * pv.dx - 0.5 * pv.dx;
        double tright = pv.x1 + (double)(iright - 1) * pv.dx + 0.5 * pv.dx;
        if (tleft < pv.xmin) tleft = pv.xmin;
        if (tright > pv.xmax) tright = pv.xmax;

        // raw output region of this stretch
        const double cprime = ps.cprime;
        const long long region = (long long)ps.cap_start[clip] + (long long)floor(tleft * cprime) + 8LL * k;
        const int cap = (int)floor((tright - tleft) * cprime) + 8;
        double* rt = ps.raw_t + 2 * region + (go_right ? cap : 0);
        double* rthr = ps.raw_thr + 2 * region;
        int nl = 0, nr = 0;
        double addedRight = -1e308;

        const double tmiddle = (tleft + tright) / 2;
        const double f0middle = pitch_value_at(pv, tmiddle);
        if (!is_undef(f0middle)) {
            double tmax = find_extremum_warp(S, tmiddle - 0.5 / f0middle, tmiddle + 0.5 / f0middle, lane);
            const double tsave = tmax;
            double peak;
            if (!go_right) {
                for (;;) {
                    double f0 = pitch_value_at(pv, tmax);
                    if (is_undef(f0)) break;
                    double tnew = tmax;
                    double correlation = find_max_corr_warp(S, tmax, 1.0 / f0, tmax - 1.25 / f0, tmax - 0.8 / f0, &tnew, &peak, lane);
                    tmax = tnew;
                    if (correlation == -1.0) tmax -= 1.0 / f0;
                    if (tmax < tleft) {
                        if (correlation > 0.7 && peak > 0.023333 * globalPeak) {
                            if (nl < cap) { if (lane == 0) { rt[nl] = tmax; rthr[nl] = 0.8 / f0; } nl++; }
                        }
                        break;
                    }
                    if (correlation > 0.3 && (peak == 0.0 || peak > 0.01 * globalPeak)) {
                        if (nl < cap) { if (lane == 0) { rt[nl] = tmax; rthr[nl] = 0.8 / f0; } nl++; }
                    }
                }
            } else {
                // first point, then walk right
                if (lane == 0) rt[0] = tsave;
                nr = 1;
                for (;;) {
                    double f0 = pitch_value_at(pv, tmax);
                    if (is_undef(f0)) break;
                    double tnew = tmax;
                    double correlation = find_max_corr_warp(S, tmax, 1.0 / f0, tmax + 0.8 / f0, tmax + 1.25 / f0, &tnew, &peak, lane);
                    tmax = tnew;
                    if (correlation == -1.0) tmax += 1.0 / f0;
                    if (tmax > tright) {
                        if (correlation > 0.7 && peak > 0.023333 * globalPeak) {
                            if (nr < cap) { if (lane == 0) rt[nr] = tmax; nr++; addedRight = tmax; }
                        }
                        break;
                    }
                    if (correlation > 0.3 && (peak == 0.0 || peak > 0.01 * globalPeak)) {
                        if (nr < cap) { if (lane == 0) rt[nr] = tmax; nr++; addedRight = tmax; }
                    }
                }
            }
        }
        if (lane == 0) {
            if (!go_right) {
                ps.raw_nleft[slot] = nl;
                ps.raw_region[slot] = 2 * region + ((long long)cap << 40);
            } else {
                ps.raw_nright[slot] = nr;
                ps.raw_added_right[slot] = addedRight;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ per-clip assembly
// Applies the `tmax - addedRight > 0.8/f0` guard of the left walks (addedRight comes from earlier stretches only),
// concatenates the stretches and restores PointProcess_addPoint's sorted order.
__global__ void k_pulses_assemble(Clips c, PitchPass p, PulseSet ps) {
    const int clip = blockIdx.x * blockDim.x + threadIdx.x;
    if (clip >= c.n) return;
    const int ns = ps.st_count[clip];
    const int f0 = p.fstart[clip];
    double* out = ps.t + ps.cap_start[clip];
    int n = 0;
    double addedRight = -1e308;
    for (int k = 0; k < ns; k++) {
        const int slot = f0 + k;
        const long long packed = ps.raw_region[slot];
        const long long region2 = packed & ((1LL << 40) - 1);
        const int cap = (int)(packed >> 40);
        const double* rt = ps.raw_t + region2;
        const double* rthr = ps.raw_thr + region2;
        const int nl = ps.raw_nleft[slot], nr = ps.raw_nright[slot];
        const int first = n;
        for (int i = nl - 1; i >= 0; i--)
            if (rt[i] - addedRight > rthr[i]) out[n++] = rt[i];
        for (int i = 0; i < nr; i++) out[n++] = rt[cap + i];
        double ar = ps.raw_added_right[slot];
        if (ar != -1e308) addedRight = ar;
        // sorted insertion of the new points (only the seam with the previous stretch can be out of order)
        for (int i = first; i < n; i++) {
            double v = out[i];
            int j = i;
            while (j > 0 && out[j - 1] > v) { out[j] = out[j - 1]; j--; }
            out[j] = v;
        }
    }
    ps.count[clip] = n;
}

void launch_pulses(const Clips& c, const PitchPass& p, const PulseSet& ps, cudaStream_t s) {
    k_stretch_list<<<(c.n + 3) / 4, 128, 0, s>>>(c, p, ps);
    launch_exclusive_scan(ps.st_count, ps.st_start, c.n, s);
    int grid = sm_count() * 8;
    k_pulses_stretch<<<grid, PW * 32, 0, s>>>(c, p, ps);
    k_pulses_assemble<<<(c.n + 63) / 64, 64, 0, s>>>(c, p, ps);
}
