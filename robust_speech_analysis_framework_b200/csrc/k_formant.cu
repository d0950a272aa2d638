// k_formant.cu -- Sound_to_Formant_burg (fon/Sound_to_Formant.cpp): pre-emphasis, Gaussian window, Burg LPC (dwsys/NUM2
// VECburg), polynomial roots (Praat: companion-matrix eigenvalues + Newton polish, dwsys/Roots.cpp), unit-circle fix,
// frequency / bandwidth extraction and sorting -- and the formant queries of _measureFormants at the glottal pulses
// (mshds_extractor.py:319-336).
//
// One warp per 5 ms frame of the 10 kHz signal: the 500-sample lattice vectors b1/b2 live in shared memory, the ten
// reflection steps are warp reductions, and the ten roots are found by lanes 0..9 with the Aberth-Ehrlich simultaneous
// iteration (a different algorithm from the oracle's Hessenberg QR: both converge to the same roots) + Newton polish.
#include "internal.h"
#include "common.cuh"

#define FW 4                   // warps per CTA
#define FWIN_MAX 512           // >= nsamp_window (500 at 10 kHz)
#define NPOLES 10

__global__ void k_formant_grid(FormantPass p, int njobs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= njobs) return;
    const ResampleJob J = p.jobs[i];
    const double dx = J.out_dx;
    const double duration = (double)J.nout * dx;
    int nFrames = 1 + (int)floor((duration - p.dt_window) / p.dt);
    double t1 = J.out_x1 + 0.5 * (duration - dx - (double)(nFrames - 1) * p.dt);
    if (nFrames < 1 || p.nsamp_window > J.nout) nFrames = 0;      // shorter than one window: handled as failure
    p.nF[i] = nFrames;
    p.t1[i] = t1;
}

struct cplx { double re, im; };
__device__ __forceinline__ cplx c_mul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx c_div(cplx a, cplx b) {
    double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
__device__ __forceinline__ double c_abs(cplx a) { return hypot(a.re, a.im); }
// p(z) and p'(z) for coefficients c[0..n] in ascending powers
__device__ __forceinline__ void poly_eval(const double* c, int n, cplx z, cplx* p, cplx* dp) {
    cplx pv = {c[n], 0.0}, dv = {0.0, 0.0};
    for (int i = n - 1; i >= 0; i--) {
        dv = c_mul(dv, z); dv.re += pv.re; dv.im += pv.im;
        pv = c_mul(pv, z); pv.re += c[i];
    }
    *p = pv; *dp = dv;
}

__device__ __forceinline__ cplx c_rcp(cplx b) {                       // 1 / b with a single division
    double inv = 1.0 / (b.re * b.re + b.im * b.im);
    return {b.re * inv, -b.im * inv};
}

#define FGROUP 3               // frames per warp pass: their 3 x 10 roots are iterated together on lanes 0..29

__global__ void __launch_bounds__(FW * 32) k_formant_frames(FormantPass p, int njobs, const double* __restrict__ sig) {
    __shared__ double s_b1[FW][FWIN_MAX], s_b2[FW][FWIN_MAX];
    __shared__ double s_coef[FW][FGROUP][NPOLES + 2];
    __shared__ int s_ok[FW][FGROUP];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = blockIdx.x * FW + wib, nw = gridDim.x * FW;
    const int total = p.fstart[njobs];
    const int ngroups = (total + FGROUP - 1) / FGROUP;
    double* b1 = s_b1[wib]; double* b2 = s_b2[wib];
    const int n = p.nsamp_window, half = p.nsamp_window / 2;
    for (int grp = gw; grp < ngroups; grp += nw) {
        // ---- Burg (VECburg) for up to FGROUP consecutive frames, one after the other on the full warp
        for (int sub = 0; sub < FGROUP; sub++) {
            const int f = grp * FGROUP + sub;
            double* coef = s_coef[wib][sub];
            int okf = 0;
            if (f < total) {
                const int job = find_segment(p.fstart, njobs, f);
                const ResampleJob J = p.jobs[job];
                const double* y = sig + J.out_off - 1;                  // 1-based resampled sound (not yet pre-emphasised)
                const long long nx = J.nout;
                const double dx = J.out_dx, x1 = J.out_x1;
                const double t = p.t1[job] + (double)(f - p.fstart[job]) * p.dt;
                const long long leftSample = x_to_low(x1, dx, t), rightSample = leftSample + 1;
                long long startSample = rightSample - half, endSample = leftSample + half;
                if (startSample < 1) startSample = 1;
                if (endSample > nx) endSample = nx;
                // Sound_preEmphasis on the fly: s'[i] = s[i] - e*s[i-1] (i >= 2), s'[1] = s[1]
                double maxI = 0.0;
                for (long long i = startSample + lane; i <= endSample; i += 32) {
                    double v = i >= 2 ? y[i] - p.emphasis * y[i - 1] : y[i];
                    maxI = fmax(maxI, v * v);
                }
                maxI = warp_max(maxI);
                if (maxI != 0.0) {
                    __syncwarp();
                    // windowed frame x[1..n] -> lattice vectors (b1[j] = x[j], j<n ; b2[j-1] = x[j], j>1), 0-based storage
                    double pacc = 0.0;
                    for (int j = 1 + lane; j <= n; j += 32) {
                        long long i = startSample + j - 1;
                        double v = 0.0;
                        if (i <= nx) v = (i >= 2 ? y[i] - p.emphasis * y[i - 1] : y[i]) * __ldg(p.window + j - 1);
                        pacc += v * v;
                        if (j <= n - 1) b1[j - 1] = v;
                        if (j >= 2) b2[j - 2] = v;
                    }
                    pacc = warp_sum(pacc);
                    __syncwarp();
                    double a[NPOLES + 1], aa[NPOLES + 1];
                    for (int j = 0; j <= NPOLES; j++) { a[j] = 0.0; aa[j] = 0.0; }
                    const bool okb = pacc / (double)n > 0.0;
                    for (int i = 1; i <= NPOLES && okb; i++) {
                        double num = 0.0, den = 0.0;
                        for (int j = lane; j < n - i; j += 32) {
                            double u = b1[j], w = b2[j];
                            num += u * w;
                            den += u * u + w * w;
                        }
                        num = warp_sum(num); den = warp_sum(den);
                        if (den <= 0.0) break;                      // VECburg returns with the coefficients found so far
                        a[i] = 2.0 * (num / den);
                        for (int j = 1; j <= i - 1; j++) a[j] = aa[j] - a[i] * aa[i - j];
                        if (i < NPOLES) {
                            for (int j = 1; j <= i; j++) aa[j] = a[j];
                            const double ai = aa[i];
                            // b1[j] -= ai*b2[j]; b2[j] = b2[j+1] - ai*b1[j+1]   (right-hand sides are the OLD values)
                            for (int j0 = 0; j0 < n - i - 1; j0 += 32) {
                                int j = j0 + lane;
                                double nb1 = 0.0, nb2 = 0.0;
                                bool act = j < n - i - 1;
                                if (act) { nb1 = b1[j] - ai * b2[j]; nb2 = b2[j + 1] - ai * b1[j + 1]; }
                                __syncwarp();
                                if (act) { b1[j] = nb1; b2[j] = nb2; }
                                __syncwarp();
                            }
                        }
                    }
                    // polynomial z^10 - sum a_k z^(10-k): ascending coefficients c[i-1] = -a[10-i+1], c[10] = 1
                    if (lane == 0) {
                        for (int i = 1; i <= NPOLES; i++) coef[i - 1] = -a[NPOLES - i + 1];
                        coef[NPOLES] = 1.0;
                    }
                    okf = okb ? 1 : 0;
                }
            }
            if (lane == 0) s_ok[wib][sub] = okf;
            __syncwarp();
        }
        // ---- roots of the FGROUP polynomials together: lane = 10 * sub + root
        const int sub = lane / NPOLES, ridx = lane % NPOLES;
        const bool own = lane < FGROUP * NPOLES && s_ok[wib][sub < FGROUP ? sub : 0] != 0;
        double c[NPOLES + 1];
        for (int i = 0; i <= NPOLES; i++) c[i] = own ? s_coef[wib][sub][i] : (i == NPOLES ? 1.0 : 0.5);
        double ang = 2.0 * MSHDS_PI * (double)ridx / (double)NPOLES + 0.4;
        double rad = pow(fabs(c[0]) > 1e-300 ? fabs(c[0]) : 1e-300, 1.0 / NPOLES);     // geometric mean of |roots|
        if (rad < 0.3) rad = 0.3;
        if (rad > 1.2) rad = 1.2;
        cplx z = {rad * cos(ang), rad * sin(ang)};
        const int base = (lane < FGROUP * NPOLES ? sub : FGROUP - 1) * NPOLES;
        // Aberth-Ehrlich simultaneous iteration; a frame whose own roots have converged is frozen, so its result never
        // depends on the frames it happens to share the warp with (batch-composition invariance)
        bool frozen = !own;
        for (int it = 0; it < 60; it++) {
            cplx pz, dpz;
            poly_eval(c, NPOLES, z, &pz, &dpz);
            const bool zero_p = (pz.re == 0.0 && pz.im == 0.0);
            cplx newton = c_mul(pz, c_rcp(dpz));
            cplx sum = {0.0, 0.0};
            for (int k = 0; k < NPOLES; k++) {
                double zr = __shfl_sync(FULL_MASK, z.re, base + k), zi = __shfl_sync(FULL_MASK, z.im, base + k);
                if (k != ridx) {
                    cplx inv = c_rcp({z.re - zr, z.im - zi});
                    sum.re += inv.re; sum.im += inv.im;
                }
            }
            cplx denom = c_mul(newton, sum);
            denom.re = 1.0 - denom.re; denom.im = -denom.im;
            cplx step = zero_p ? cplx{0.0, 0.0} : c_mul(newton, c_rcp(denom));
            if (!frozen) { z.re -= step.re; z.im -= step.im; }
            double sz = !frozen ? sqrt((step.re * step.re + step.im * step.im) / fmax(z.re * z.re + z.im * z.im, 1e-300)) : 0.0;
            if (!(sz == sz)) sz = 0.0;
            bool any_active = false;
            for (int sb = 0; sb < FGROUP; sb++) {
                double m = warp_max(sub == sb ? sz : 0.0);
                if (m < 1e-13) { if (sub == sb) frozen = true; }      // the Newton polish below finishes the last digits
                else any_active = true;
            }
            if (!any_active) break;
        }
        // Newton polish on the original polynomial (Roots_Polynomial_polish): keep the iterate with the smallest |p|
        {
            cplx pz, dpz;
            poly_eval(c, NPOLES, z, &pz, &dpz);
            double best = pz.re * pz.re + pz.im * pz.im;
            for (int it = 0; it < 80; it++) {
                if (dpz.re == 0.0 && dpz.im == 0.0) break;
                cplx q = c_mul(pz, c_rcp(dpz));
                cplx zn = {z.re - q.re, z.im - q.im};
                cplx pn, dpn;
                poly_eval(c, NPOLES, zn, &pn, &dpn);
                double fa = pn.re * pn.re + pn.im * pn.im;
                if (!(fa < best)) break;
                best = fa; z = zn; pz = pn; dpz = dpn;
            }
        }
        // Roots_fixIntoUnitCircle, then frequencies / bandwidths of the roots with Im >= 0 inside the safety margins
        double re = z.re, im = z.im;
        double a2 = re * re + im * im;
        if (a2 > 1.0) { re /= a2; im /= a2; }
        double fr = -1.0, bw = 0.0;
        if (own && im >= 0.0) {
            double fq = fabs(atan2(im, re)) * p.nyquist / MSHDS_PI;
            if (fq >= 50.0 && fq <= p.nyquist - 50.0) { fr = fq; bw = -log(re * re + im * im) * p.nyquist / MSHDS_PI; }
        }
        const unsigned mall = __ballot_sync(FULL_MASK, fr >= 0.0);
        for (int sb = 0; sb < FGROUP; sb++) {
            const int f = grp * FGROUP + sb;
            if (f >= total) break;
            // gather and sort by frequency (<= 5 formants)
            unsigned m = (mall >> (sb * NPOLES)) & ((1u << NPOLES) - 1u);
            int nform = 0;
            double ff[5], fb[5];
            while (m && nform < 5) {
                int src = __ffs(m) - 1 + sb * NPOLES;
                m &= m - 1;
                ff[nform] = __shfl_sync(FULL_MASK, fr, src);
                fb[nform] = __shfl_sync(FULL_MASK, bw, src);
                nform++;
            }
            for (int i = 1; i < nform; i++)
                for (int j = i; j > 0 && ff[j] < ff[j - 1]; j--) {
                    double tf = ff[j]; ff[j] = ff[j - 1]; ff[j - 1] = tf;
                    tf = fb[j]; fb[j] = fb[j - 1]; fb[j - 1] = tf;
                }
            if (lane == 0) {
                p.nform[f] = nform;
                for (int k = 0; k < 5; k++) {
                    p.freq[(size_t)f * 5 + k] = k < nform ? ff[k] : DEVNAN;
                    p.bw[(size_t)f * 5 + k] = k < nform ? fb[k] : DEVNAN;
                }
            }
        }
        __syncwarp();
    }
}

// Formant "Get value at time" / "Get bandwidth at time" (Hertz, Linear) = Sampled_getValueAtX on the frame grid
__device__ double formant_value_at(const FormantPass& p, int job, double xmax, int iformant, double x, bool bandwidth) {
    if (x < 0.0 || x > xmax) return DEVNAN;
    const int nx = p.nF[job], f0 = p.fstart[job];
    double ireal = (x - p.t1[job]) / p.dt + 1.0;
    long long ileft = (long long)floor(ireal), inear, ifar;
    double phase = ireal - (double)ileft;
    if (phase < 0.5) { inear = ileft; ifar = ileft + 1; }
    else { ifar = ileft; inear = ileft + 1; phase = 1.0 - phase; }
    if (inear < 1 || inear > nx) return DEVNAN;
    if (iformant > p.nform[f0 + inear - 1]) return DEVNAN;
    const double* arr = bandwidth ? p.bw : p.freq;
    double fnear = arr[(size_t)(f0 + inear - 1) * 5 + iformant - 1];
    if (ifar < 1 || ifar > nx) return fnear;
    if (iformant > p.nform[f0 + ifar - 1]) return fnear;
    double ffar = arr[(size_t)(f0 + ifar - 1) * 5 + iformant - 1];
    return fnear + phase * (ffar - fnear);
}

// mean / std(ddof=1) of F1, B1, F2, B2 sampled at the glottal pulses (mshds_extractor.py:324-336); job index == clip index
__global__ void __launch_bounds__(256) k_formant_stats(Clips c, FormantPass p, PulseSet ps) {
    __shared__ double red[32];
    const int clip = blockIdx.x;
    const double* t = ps.t + ps.cap_start[clip];
    const int np = ps.count[clip];
    const double xmax = c.xmax[clip];
    double* feat = c.feat + (size_t)clip * N_FEAT + 13;
    const bool valid = p.nF[clip] >= 1 && ps.valid[clip];
    for (int q = 0; q < 4; q++) {
        const int iformant = q < 2 ? 1 : 2;
        const bool bandwidth = (q & 1) != 0;
        double s = 0.0, n = 0.0;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            double v = formant_value_at(p, clip, xmax, iformant, t[i], bandwidth);
            if (!is_undef(v)) { s += v; n += 1.0; }
        }
        s = block_sum(s, red); n = block_sum(n, red);
        double mean = DEVNAN, sd = DEVNAN;
        if (valid && n > 0.0) {
            mean = s / n;
            if (n > 1.0) {
                double v2 = 0.0;
                for (int i = threadIdx.x; i < np; i += blockDim.x) {
                    double v = formant_value_at(p, clip, xmax, iformant, t[i], bandwidth);
                    if (!is_undef(v)) v2 += (v - mean) * (v - mean);
                }
                v2 = block_sum(v2, red);
                sd = sqrt(v2 / (n - 1.0));
            }
        }
        if (threadIdx.x == 0) {
            feat[2 * q] = mean; feat[2 * q + 1] = sd;
            // status bit = "mean_F1_Loc is NaN" (helper raised, or no pulse produced a value)
            if (q == 0 && is_undef(mean)) atomicOr(&c.status[clip], ST_FORMANT);
        }
    }
}

void launch_formants(const Clips& c, const FormantPass& p, int njobs, const double* sig, int max_frames_hint, cudaStream_t s) {
    k_formant_grid<<<(njobs + 127) / 128, 128, 0, s>>>(p, njobs);
    launch_exclusive_scan(p.nF, p.fstart, njobs, s);
    int grid = (max_frames_hint / FGROUP + FW) / FW;
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    if (grid < 1) grid = 1;
    k_formant_frames<<<grid, FW * 32, 0, s>>>(p, njobs, sig);
}
void launch_formant_stats(const Clips& c, const FormantPass& p, const PulseSet& ps, cudaStream_t s) {
    k_formant_stats<<<c.n, 256, 0, s>>>(c, p, ps);
}
