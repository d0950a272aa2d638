// k_ltas.cu -- PointProcess_Sound_to_Ltas (fon/Ltas.cpp), Ltas_getSlope and the robust (Theil) tilt line of
// dwtools/Ltas_extensions.cpp: "To Ltas (pitch-corrected)... floor ceiling 5000 100 0.0001 0.02 1.3",
// "Get slope 50 1000 1000 4000 dB" and "Report spectral tilt 100 5000 Linear Robust" (mshds_extractor.py:241-248).
//
// Every glottal period (pulse with two admissible neighbouring intervals) is an exact-length DFT of 32..320 samples up
// to 5 kHz.  One warp owns LTAS_PART consecutive pulses of one clip: samples and the cos/sin table of the period are
// staged in shared memory, lanes own DFT bins (each bin summed in sample order like a serial loop), band sums are
// formed in bin order and parts are combined in pulse order, so the result is deterministic.
#include "internal.h"
#include "common.cuh"

#define LTAS_PART 64
#define LW 4                 // warps per CTA
#define NMAX 352             // longest period in samples (0.02 s at 16 kHz = 320) + slack
#define NBAND 50

__global__ void k_ltas_parts(Clips c, PulseSet ps, LtasPass lt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    int np = ps.count[i] - 2;                       // candidate periods: pulses 2 .. nt-1
    lt.part_count[i] = np > 0 ? (np + LTAS_PART - 1) / LTAS_PART : 0;
    lt.fail[i] = np < 1 ? 1 : 0;
}

__global__ void __launch_bounds__(LW * 32) k_ltas_accum(Clips c, PulseSet ps, LtasPass lt, double maximumFrequency,
                                                         double bandWidth, double shortestPeriod, double longestPeriod,
                                                         double maximumPeriodFactor) {
    __shared__ double s_seg[LW][NMAX], s_cos[LW][NMAX], s_sin[LW][NMAX], s_e[LW][NMAX / 2 + 8];
    __shared__ double s_band[LW][2 * NBAND];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * LW + wib, nw = gridDim.x * LW;
    const int total = lt.part_start[c.n];
    double* seg = s_seg[wib]; double* ct = s_cos[wib]; double* st = s_sin[wib]; double* eb = s_e[wib]; double* band = s_band[wib];
    for (int w = gw; w < total; w += nw) {
        const int clip = find_segment(lt.part_start, c.n, w);
        const int part = w - lt.part_start[clip];
        const double* t1b = ps.t + ps.cap_start[clip] - 1;          // 1-based pulses
        const int nt = ps.count[clip];
        const long long nx = c.off[clip + 1] - c.off[clip];
        const SPtr pcm = c.pcm + c.off[clip];
        const double dx = c.dx, x1 = c.x1[clip];
        for (int b = lane; b < 2 * NBAND; b += 32) band[b] = 0.0;
        __syncwarp();
        int ip0 = 2 + part * LTAS_PART, ip1 = ip0 + LTAS_PART - 1;
        if (ip1 > nt - 1) ip1 = nt - 1;
        for (int ipulse = ip0; ipulse <= ip1; ipulse++) {
            double leftInterval = t1b[ipulse] - t1b[ipulse - 1];
            double rightInterval = t1b[ipulse + 1] - t1b[ipulse];
            double intervalFactor = leftInterval > rightInterval ? leftInterval / rightInterval : rightInterval / leftInterval;
            if (!(leftInterval >= shortestPeriod && leftInterval <= longestPeriod && rightInterval >= shortestPeriod &&
                  rightInterval <= longestPeriod && intervalFactor <= maximumPeriodFactor))
                continue;
            double ta = t1b[ipulse] - 0.5 * leftInterval, tb = t1b[ipulse] + 0.5 * rightInterval;
            long long ix1 = 1 + (long long)ceil((ta - x1) / dx);
            long long ix2 = 1 + (long long)floor((tb - x1) / dx);
            if (ix2 < ix1) { if (lane == 0) lt.fail[clip] = 1; continue; }       // "would contain no samples"
            int n = (int)(ix2 - ix1 + 1);
            if (n > NMAX) n = NMAX;                                               // cannot happen: period <= 0.02 s
            for (int j = lane; j < n; j += 32) {
                long long is = ix1 + j;
                seg[j] = (is >= 1 && is <= nx) ? samp(pcm, is - 1) : 0.0;
                double ang = 2.0 * MSHDS_PI * (double)j / (double)n;
                ct[j] = cos(ang); st[j] = sin(ang);
            }
            __syncwarp();
            const int nfreq = n / 2 + 1;
            const double df = 1.0 / (dx * n);
            for (int ifreq = 1 + lane; ifreq <= nfreq; ifreq += 32) {
                double frequency = (ifreq - 1) * df;
                long long iband = (long long)ceil(frequency / bandWidth);
                double energy = -1.0;
                if (iband >= 1 && iband <= NBAND) {
                    double re = 0.0, im = 0.0;
                    int ph = 0;
                    for (int j = 0; j < n; j++) {
                        re += seg[j] * ct[ph];
                        im -= seg[j] * st[ph];
                        ph += ifreq - 1;
                        if (ph >= n) ph -= n;
                    }
                    re *= dx; im *= dx;
                    if (ifreq == nfreq && (n & 1) == 0) im = 0.0;
                    energy = (re * re + im * im) * 2.0 * df;
                }
                eb[ifreq] = energy;
            }
            __syncwarp();
            for (int b = lane; b < NBAND; b += 32) {
                double e = band[b], cnt = band[NBAND + b];
                for (int ifreq = 1; ifreq <= nfreq; ifreq++) {
                    double frequency = (ifreq - 1) * df;
                    if ((long long)ceil(frequency / bandWidth) == b + 1) { e += eb[ifreq]; cnt += 1.0; }
                }
                band[b] = e; band[NBAND + b] = cnt;
            }
            __syncwarp();
        }
        for (int b = lane; b < 2 * NBAND; b += 32) lt.partial[(size_t)w * 2 * NBAND + b] = band[b];
        __syncwarp();
    }
}

// dwsys/NUM2 NUMlineFit_theil, incomplete variant, on small arrays (single thread)
__device__ void sort_small(double* a, int n) {
    for (int i = 1; i < n; i++) { double v = a[i]; int j = i; while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; j--; } a[j] = v; }
}
__device__ double quantile_small(const double* a0, int n, double factor) {   // a0 sorted, 0-based
    double place = factor * n + 0.5;
    int left = (int)floor(place);
    if (n < 1) return 0.0;
    if (left < 1) return a0[0];
    if (left >= n) return a0[n - 1];
    if (a0[left] == a0[left - 1]) return a0[left - 1];
    return a0[left - 1] + (place - left) * (a0[left] - a0[left - 1]);
}

// Sampled_getMean (interpolate = false) on the Ltas in dB, as used by Ltas_getSlope
__device__ double ltas_mean_dB(const double* z0, int nx, double dxb, double xmin, double xmax) {
    double x1 = 0.5 * dxb;
    double sum = 0.0, definitionRange = 0.0;
    double dom_max = nx * dxb;
    if (xmax <= xmin) { xmin = 0.0; xmax = dom_max; }
    if (xmin < 0.0) xmin = 0.0;
    if (xmax > dom_max) xmax = dom_max;
    if (xmin >= xmax) return DEVNAN;
    double rimin = (xmin - x1) / dxb + 1.0, rimax = (xmax - x1) / dxb + 1.0;
    if (rimax >= 0.5 && rimin < nx + 0.5) {
        int imin = rimin < 0.5 ? 0 : (int)iround_d(rimin);
        int imax = rimax >= nx + 0.5 ? nx + 1 : (int)iround_d(rimax);
        for (int isamp = imin + 1; isamp < imax; isamp++) { definitionRange += 1.0; sum += z0[isamp - 1]; }
        if (imin == imax) {
            if (imin >= 1 && imin <= nx) { double phase = rimax - rimin; definitionRange += phase; sum += phase * z0[imin - 1]; }
        } else {
            if (imin >= 1) { double phase = imin - rimin + 0.5; definitionRange += phase; sum += phase * z0[imin - 1]; }
            if (imax <= nx) { double phase = rimax - imax + 0.5; definitionRange += phase; sum += phase * z0[imax - 1]; }
        }
    }
    if (definitionRange <= 0.0) return DEVNAN;
    return sum / definitionRange;
}

__global__ void k_ltas_final(Clips c, PulseSet ps, LtasPass lt, double* ltas_bands, double bandWidth) {
    const int clip = blockIdx.x * blockDim.x + threadIdx.x;
    if (clip >= c.n) return;
    double* z = ltas_bands + (size_t)clip * NBAND;
    bool fail = lt.fail[clip] != 0;
    double e[NBAND], cnt[NBAND];
    for (int b = 0; b < NBAND; b++) { e[b] = 0.0; cnt[b] = 0.0; }
    const int p0 = lt.part_start[clip], np = lt.part_count[clip];
    for (int p = 0; p < np; p++)
        for (int b = 0; b < NBAND; b++) {
            e[b] += lt.partial[(size_t)(p0 + p) * 2 * NBAND + b];
            cnt[b] += lt.partial[(size_t)(p0 + p) * 2 * NBAND + NBAND + b];
        }
    double totalNumberOfEnergies = 0.0;
    for (int b = 0; b < NBAND; b++) totalNumberOfEnergies += cnt[b];
    if (totalNumberOfEnergies == 0.0) fail = true;          // no valid period / no energy in any bin
    const double duration = c.xmax[clip];                     // sound -> xmax - sound -> xmin
    if (!fail) {
        for (int b = 0; b < NBAND; b++) {
            if (cnt[b] == 0.0) z[b] = DEVNAN;
            else {
                double meanEnergyInThisBand = e[b] / cnt[b];
                double meanNumberOfEnergiesPerBand = totalNumberOfEnergies / NBAND;
                double redistributedEnergyInThisBand = meanEnergyInThisBand * meanNumberOfEnergiesPerBand;
                double redistributedEnergyDensityInThisBand = redistributedEnergyInThisBand / bandWidth;
                double redistributedPowerDensityInThisBand = redistributedEnergyDensityInThisBand / duration;
                z[b] = 10.0 * log10(redistributedPowerDensityInThisBand / 4.0e-10);
            }
        }
        const double x1 = 0.5 * bandWidth;
        for (int iband = 1; iband <= NBAND; iband++) {
            if (is_undef(z[iband - 1])) {
                int ibandleft = iband - 1, ibandright = iband + 1;
                while (ibandleft >= 1 && is_undef(z[ibandleft - 1])) ibandleft--;
                while (ibandright <= NBAND && is_undef(z[ibandright - 1])) ibandright++;
                if (ibandleft < 1 && ibandright > NBAND) { fail = true; break; }
                if (ibandleft < 1) z[iband - 1] = z[ibandright - 1];
                else if (ibandright > NBAND) z[iband - 1] = z[ibandleft - 1];
                else {
                    double frequency = x1 + (iband - 1) * bandWidth;
                    double fleft = x1 + (ibandleft - 1) * bandWidth;
                    double fright = x1 + (ibandright - 1) * bandWidth;
                    z[iband - 1] = ((fright - frequency) * z[ibandleft - 1] + (frequency - fleft) * z[ibandright - 1]) / (fright - fleft);
                }
            }
        }
    }
    double slope = DEVNAN, tilt = DEVNAN;
    if (!fail) {
        double low = ltas_mean_dB(z, NBAND, bandWidth, 50.0, 1000.0);
        double high = ltas_mean_dB(z, NBAND, bandWidth, 1000.0, 4000.0);
        if (!is_undef(low) && !is_undef(high)) slope = high - low;
        // Ltas_fitTiltLine 100..5000 Hz, linear frequency, Theil incomplete
        long long ifmin, ifmax;
        long long n = get_window_samples(0.5 * bandWidth, bandWidth, NBAND, 100.0, 5000.0, &ifmin, &ifmax);
        if (n >= 2) {
            double x[NBAND], y[NBAND], mbs[NBAND];
            for (long long i = ifmin; i <= ifmax; i++) { x[i - ifmin] = 0.5 * bandWidth + (i - 1) * bandWidth; y[i - ifmin] = z[i - 1]; }
            int nn = (int)n;
            if (nn == 2) tilt = (y[1] - y[0]) / (x[1] - x[0]);
            else {
                int numberOfPairs = nn / 2;
                int n2 = (nn % 2 == 1) ? numberOfPairs + 1 : numberOfPairs;
                for (int i = 1; i <= numberOfPairs; i++) { int i2 = n2 + i; mbs[i - 1] = (y[i2 - 1] - y[i - 1]) / (x[i2 - 1] - x[i - 1]); }
                sort_small(mbs, numberOfPairs);
                tilt = quantile_small(mbs, numberOfPairs, 0.5);
            }
        } else fail = true;
    }
    if (fail) { slope = DEVNAN; tilt = DEVNAN; atomicOr(&c.status[clip], ST_LTAS); for (int b = 0; b < NBAND; b++) z[b] = DEVNAN; }
    c.feat[(size_t)clip * N_FEAT + 10] = slope;
    c.feat[(size_t)clip * N_FEAT + 11] = tilt;
}

void launch_ltas(const Clips& c, const PulseSet& ps, const LtasPass& lt, double* ltas_bands, cudaStream_t s) {
    k_ltas_parts<<<(c.n + 127) / 128, 128, 0, s>>>(c, ps, lt);
    launch_exclusive_scan(lt.part_count, lt.part_start, c.n, s);
    k_ltas_accum<<<sm_count() * 4, LW * 32, 0, s>>>(c, ps, lt, 5000.0, 100.0, 0.0001, 0.02, 1.3);
    k_ltas_final<<<(c.n + 63) / 64, 64, 0, s>>>(c, ps, lt, ltas_bands, 100.0);
}
