// k_cpp.cu -- _extract_CPP (mshds_extractor.py:253-301): PointProcess_to_TextGrid_vuv (0.02, 0.1) + "Down to Table"
// (6-decimal times), per voiced segment Sound_to_PowerCepstrogram (60, 0.002, 5000, 50) and PowerCepstrogram "Get CPPS"
// (no, 0.01, 0.001, 60, 330, 0.05, parabolic, 0.001, 0, Straight, Robust); keep segments with CPPS > 4; mean.
//
// Praat sources restated: fon/PointProcess.cpp (vuv), LPC/Sound_and_PowerCepstrogram.cpp, LPC/PowerCepstrogram.cpp
// (smooth, getCPPS), LPC/PowerCepstrum.cpp (fitTiltLine, getPeakProminence), dwsys/NUM2.cpp (Theil line fit).
//
// One CTA per cepstrogram frame: mean removal, Gaussian window, packed real FFT-1024, log power and the inverse
// transform stay in shared memory; a second kernel (one WARP per frame, k_cpp_frames_warp) does the 5-frame / 10-bin box
// smoothing, the dB conversion, the robust tilt line (two medians by bitonic sorting networks held in registers) and the
// parabolic peak per frame.
#include <cstdlib>
#include "internal.h"
#define NT_CEP_DEFAULT 128
#include "common.cuh"
#include "fft.cuh"
#include "num.cuh"
#include "fftwarp.cuh"

// ------------------------------------------------------------------------------------------------ V segments
__global__ void k_vuv_segments(Clips c, PulseSet ps, CppSegs sg, double maxT, double meanT) {
    const int clip = blockIdx.x * blockDim.x + threadIdx.x;
    if (clip >= c.n) return;
    const double* t1b = ps.t + ps.cap_start[clip] - 1;           // 1-based pulses
    const int nt = ps.count[clip];
    const double dx = c.dx, x1 = c.x1[clip], xmin = 0.0, xmax = c.xmax[clip];
    const double halfMeanT = 0.5 * meanT;
    const int base = sg.cap_start[clip];
    const int cap = sg.cap_start[clip + 1] - base;
    int n = 0, fail = 0;
    int ipointright;
    for (int ipointleft = 1; ipointleft <= nt; ipointleft = ipointright + 1) {
        for (ipointright = ipointleft + 1; ipointright <= nt; ipointright++)
            if (t1b[ipointright] - t1b[ipointright - 1] > maxT) break;
        ipointright--;
        double beginVoiced = t1b[ipointleft] - halfMeanT;
        if (beginVoiced < xmin) beginVoiced = xmin;
        double endVoiced = t1b[ipointright] + halfMeanT;
        if (endVoiced > xmax) endVoiced = xmax;
        // "Down to Table ... 6 decimals" then float(): round to the nearest multiple of 1e-6
        double tmin = rint(beginVoiced * 1e6) / 1e6, tmax = rint(endVoiced * 1e6) / 1e6;
        if (tmin >= tmax) continue;                                               // :284
        // Sound_extractPart (rectangular, preserve_times = False)
        long long ix1 = 1 + (long long)ceil((tmin - x1) / dx);
        long long ix2 = 1 + (long long)floor((tmax - x1) / dx);
        if (ix2 < ix1) { fail = 1; break; }
        if (n < cap) {
            sg.tmin[base + n] = tmin; sg.tmax[base + n] = tmax;
            sg.ix1[base + n] = ix1; sg.nseg[base + n] = (int)(ix2 - ix1 + 1);
            n++;
        }
    }
    sg.count[clip] = n;
    sg.fail[clip] = fail || !ps.valid[clip];
}

// ------------------------------------------------------------------------------------------------ cepstrogram frames
struct LogPowerF {
    double dx2, df;
    __device__ __forceinline__ double operator()(double p) const { return log(p * dx2 + 1e-300) * df; }
};

__global__ void __launch_bounds__(256) k_cepstrogram(const CepSeg* __restrict__ segs, const int* __restrict__ fprefix, int nsegs,
                                                      const ResampleJob* __restrict__ jobs, const double* __restrict__ sig,
                                                      const double2* __restrict__ tw, double emphasis, double dt,
                                                      double* __restrict__ cep, int nqmax, const double* __restrict__ wtab,
                                                      int wtab_n, int skip_nfft /* frames of this transform size were done by k_cepstrogram_w */) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* a = (double2*)smem;                      // 512 complex
    double* red = (double*)(smem + sizeof(double2) * 512);
    __shared__ int s_seg;
    const int total = fprefix[nsegs];
    // 8 consecutive frames per turn: neighbouring frames share most of their samples (L1 hits)
    for (int turn = blockIdx.x; turn * 8 < total; turn += gridDim.x)
    for (int f = turn * 8; f < total && f < turn * 8 + 8; f++) {
        __syncthreads();
        if (threadIdx.x == 0) s_seg = find_segment(fprefix, nsegs, f);
        __syncthreads();
        const int sgi = s_seg;
        const CepSeg S = segs[sgi];
        if (S.nfft == skip_nfft) continue;
        const ResampleJob J = jobs[sgi];
        const double* y = sig + J.out_off - 1;                                  // 1-based resampled segment
        const int iframe = f - fprefix[sgi];
        const double t = S.t1 + (double)iframe * dt;
        // Sound_into_Sound (sound, sframe, t - windowDuration / 2)
        const long long index = x_to_nearest(J.out_x1, J.out_dx, t - S.windowDuration / 2);
        const int nwin = S.nwin, nfft = S.nfft, M = nfft / 2;
        double* ar = (double*)a;
        double acc = 0.0;
        for (int i = threadIdx.x; i < nfft; i += blockDim.x) {
            double v = 0.0;
            if (i < nwin) {
                long long j = index + i;                                         // index - 1 + (i+1)
                if (j >= 1 && j <= J.nout) v = j >= 2 ? y[j] - emphasis * y[j - 1] : y[j];
            }
            ar[SWZD(i)] = v;
            acc += v;
        }
        const double mean = block_sum(acc, red) / (double)nwin;                  // Vector_subtractMean
        const double imid = 0.5 * (double)(nwin + 1), edge = exp(-12.0);
        for (int i = threadIdx.x; i < nwin; i += blockDim.x) {
            double w;
            if (nwin == wtab_n) w = __ldg(wtab + i);                 // Sound_createGaussian for the usual 0.1 s window
            else {
                double d = (double)(i + 1) - imid;
                w = (exp(-48.0 * d * d / (double)(nwin + 1) / (double)(nwin + 1)) - edge) / (1.0 - edge);
            }
            ar[SWZD(i)] = (ar[SWZD(i)] - mean) * w;
        }
        __syncthreads();
        fft_dif<-1>(a, M, tw);
        LogPowerF F;
        F.dx2 = J.out_dx * J.out_dx;
        F.df = 1.0 / (J.out_dx * (double)nfft);
        packed_power_to_inverse_input(a, M, S.logM, tw, F, (double*)nullptr);
        fft_dit<+1>(a, M, tw);
        const int nq = M + 1;
        double* row = cep + (size_t)f * nqmax;
        for (int i = threadIdx.x; i < nq; i += blockDim.x) {
            // c[i] for i <= M: natural order ar[i]; c[M] = ar[M] exists since M < nfft
            double cv = ar[SWZD(i)];
            row[i] = cv * cv;
        }
    }
}


// ------------------------------------------------------------------------------------------------ warp-per-frame version
// Round 2: half a warp per cepstrogram frame of the usual 1024-point transform (16 lanes x 32 register-resident complex
// points, fftwarp.cuh fw_roundtrip with G = log power): pre-emphasis, mean, Gaussian window, FFT, log, inverse FFT and the
// squared cepstrum without a block barrier.  Frames of shorter segments (other transform sizes) stay on k_cepstrogram.
#define CEW_WARPS 4
#define CEW_TW_BYTES (32 * 16 * 16)           // shared-memory copy of the [32][16] pass twiddles
struct CewLog {
    double dx2, df;
    __device__ __forceinline__ double operator()(double p) const { return log(p * dx2 + 1e-300) * df; }
};
__global__ void __launch_bounds__(32 * CEW_WARPS, 2) k_cepstrogram_w(const CepSeg* __restrict__ segs, const int* __restrict__ fprefix, int nsegs,
                                                                     const ResampleJob* __restrict__ jobs, const double* __restrict__ sig,
                                                                     const double2* __restrict__ tw, const double2* __restrict__ twb512,
                                                                     double emphasis, double dt, double* __restrict__ cep, int nqmax,
                                                                     const double* __restrict__ wtab, int wtab_n, int* __restrict__ turn_counter) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int L = 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = lane & 15, gidx = lane >> 4;
    const unsigned gmask = 0xffffu << (16 * gidx);
    double2* xch = (double2*)(smem + CEW_TW_BYTES + (size_t)warp * (1024 * 16) + (size_t)gidx * (512 * 16));
    fw_stage_twiddles<32 * CEW_WARPS>((double2*)smem, twb512, 16);       // [pass twiddles][exchange x warps]
    __syncthreads();
    const FwTwShared twf((const double2*)smem, j, L);
    const int total = fprefix[nsegs];
    const double2 wj = __ldg(tw + j * (TW_N / 1024));
    // a warp takes two consecutive frames per turn (neighbouring frames share 98 % of their samples: L1 hits)
    for (;;) {
        int turn = 0;
        if (lane == 0) turn = atomicAdd(turn_counter, 1);
        turn = __shfl_sync(FULL_MASK, turn, 0);
        if (2 * turn >= total) break;
        const int f = 2 * turn + gidx;
        bool active = f < total;
        int sgi = 0;
        if (active) sgi = find_segment(fprefix, nsegs, f);
        const CepSeg S = segs[sgi];
        if (S.nfft != 1024) active = false;                                      // done by k_cepstrogram
        if (__ballot_sync(FULL_MASK, active) == 0u) continue;
        const ResampleJob J = jobs[sgi];
        const double* y = sig + J.out_off - 1;                                   // 1-based resampled segment
        const int iframe = f - fprefix[sgi];
        const double t = S.t1 + (double)iframe * dt;
        const long long index = x_to_nearest(J.out_x1, J.out_dx, t - S.windowDuration / 2);    // Sound_into_Sound
        const int nwin = S.nwin;
        double2 a[32];
        double acc = 0.0;
        fr_static_for<0, 32>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const int m = 2 * (j + L * k);
            double v0 = 0.0, v1 = 0.0;
            if (active && m < nwin) {
                const long long jj = index + m;                                  // 1-based sample of frame element m
                const double ym = (jj - 1 >= 1 && jj - 1 <= J.nout) ? __ldg(y + jj - 1) : 0.0;
                const double y0 = (jj >= 1 && jj <= J.nout) ? __ldg(y + jj) : 0.0;
                const double y1 = (jj + 1 >= 1 && jj + 1 <= J.nout) ? __ldg(y + jj + 1) : 0.0;
                if (jj >= 1 && jj <= J.nout) v0 = jj >= 2 ? y0 - emphasis * ym : y0;              // Sound_preEmphasis
                if (m + 1 < nwin && jj + 1 >= 1 && jj + 1 <= J.nout) v1 = jj + 1 >= 2 ? y1 - emphasis * y0 : y1;
            }
            a[k] = make_double2(v0, v1);
            acc += v0 + v1;
        });
        const double mean = group_sum(acc, gmask, L) / (double)nwin;             // Vector_subtractMean
        const double imid = 0.5 * (double)(nwin + 1), edge = exp(-12.0);
        fr_static_for<0, 32>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const int m = 2 * (j + L * k);
            if (active && m < nwin) {
                double w0, w1;
                if (nwin == wtab_n) { w0 = __ldg(wtab + m); w1 = m + 1 < nwin ? __ldg(wtab + m + 1) : 0.0; }
                else {
                    const double d0 = (double)(m + 1) - imid, d1 = (double)(m + 2) - imid;
                    w0 = (exp(-48.0 * d0 * d0 / (double)(nwin + 1) / (double)(nwin + 1)) - edge) / (1.0 - edge);
                    w1 = (exp(-48.0 * d1 * d1 / (double)(nwin + 1) / (double)(nwin + 1)) - edge) / (1.0 - edge);
                }
                a[k].x = (a[k].x - mean) * w0;
                a[k].y = m + 1 < nwin ? (a[k].y - mean) * w1 : 0.0;
            }
        });
        CewLog G;
        G.dx2 = J.out_dx * J.out_dx;
        G.df = 1.0 / (J.out_dx * 1024.0);
        fw_roundtrip(a, xch, lane, j, L, twf, wj, G);
        // a[brev5(r)] = conj(y[n]), n = j + 16 r: cepstrum c[2n] = re, c[2n+1] = -im; the power cepstrum keeps c^2 for quefrency bins 0..512
        if (active) {
            double* row = cep + (size_t)f * nqmax;
            fr_static_for<0, 17>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const int n = j + L * r;
                const double2 v = a[fr_brev(r, 5)];
                if (2 * n <= 512) row[2 * n] = v.x * v.x;
                if (2 * n + 1 <= 512) row[2 * n + 1] = v.y * v.y;
            });
        }
    }
}

// ------------------------------------------------------------------------------------------------ CPPS per frame
// ascending bitonic sort of v[0..npow2) in shared memory (npow2 a power of two, padded with +inf by the caller).
// Pair t of a stage with partner distance j touches the 64-element window of its own warp whenever j <= 32, so those
// stages only need a warp barrier; block barriers remain for the few long-distance stages.
__device__ void block_bitonic_sort(double* v, int npow2) {
    for (int k = 2; k <= npow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (npow2 >> 1); t += blockDim.x) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // partner pairs (lo, lo + j)
                const int hi = lo + j;
                const bool up = ((lo & k) == 0);
                const double a = v[lo], b = v[hi];
                if ((a > b) == up) { v[lo] = b; v[hi] = a; }
            }
            const bool last = (k == npow2 && j == 1);
            if (j > 32 || last || (j == 1 && k >= 64)) __syncthreads();
            else __syncwarp();
        }
    }
}
// NUMquantile on a sorted 0-based array
__device__ __forceinline__ double quantile_sorted(const double* a0, int n, double factor) {
    double place = factor * n + 0.5;
    int left = (int)floor(place);
    if (n < 1) return 0.0;
    if (left < 1) return a0[0];
    if (left >= n) return a0[n - 1];
    if (a0[left] == a0[left - 1]) return a0[left - 1];
    return a0[left - 1] + (place - left) * (a0[left] - a0[left - 1]);
}

__global__ void __launch_bounds__(256) k_cpp_frames(const CepSeg* __restrict__ segs, const int* __restrict__ fprefix, int nsegs,
                                                     const double* __restrict__ cep, int nqmax, int nTimeAvg, double qAvgWindow,
                                                     double peakLo, double peakHi, double qstartFit, double qendFit,
                                                     double* __restrict__ cpp_frame) {
    __shared__ double col[520], y[520], work[1024];
    __shared__ double s_pv[8], s_px[8];
    __shared__ int s_po[8];
    __shared__ int s_seg;
    __shared__ double s_peak[2];
    const int total = fprefix[nsegs];
    for (int f = blockIdx.x; f < total; f += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) s_seg = find_segment(fprefix, nsegs, f);
        __syncthreads();
        const int sgi = s_seg;
        const CepSeg S = segs[sgi];
        const int f0 = fprefix[sgi], nFrames = fprefix[sgi + 1] - f0;
        const int i1 = f - f0 + 1;                                               // 1-based frame
        const int nq = S.nfft / 2 + 1;
        const double dq = S.dq;
        // PowerCepstrogram_smooth: moving average over time, then over quefrency
        int jfrom = i1, jto = i1;
        if (nTimeAvg > 1) {
            jfrom = i1 - nTimeAvg / 2; jto = i1 + nTimeAvg / 2;
            if ((nTimeAvg % 2) == 0) jto--;
            if (jfrom < 1) jfrom = 1;
            if (jto > nFrames) jto = nFrames;
        }
        for (int iq = threadIdx.x; iq < nq; iq += blockDim.x) {
            double s = 0.0;
            for (int j = jfrom; j <= jto; j++) s += cep[(size_t)(f0 + j - 1) * nqmax + iq];
            col[iq] = nTimeAvg > 1 ? s / (double)(jto - jfrom + 1) : s;
        }
        __syncthreads();
        const int nQ = (int)floor(qAvgWindow / dq);
        for (int iq = threadIdx.x; iq < nq; iq += blockDim.x) {
            double v;
            if (nQ > 1) {
                int i = iq + 1;
                int qf = i - nQ / 2, qt = i + nQ / 2;
                if ((nQ % 2) == 0) qt--;
                if (qf < 1) qf = 1;
                if (qt > nq) qt = nq;
                double s = 0.0;
                for (int j = qf; j <= qt; j++) s += col[j - 1];
                v = s / (double)(qt - qf + 1);
            } else v = col[iq];
            y[iq] = 10.0 * log10(v + 1e-30);                                     // PowerCepstrum dB
        }
        __syncthreads();
        // PowerCepstrum_fitTiltLine (Straight, Robust = incomplete Theil); qmax <= qmin -> whole domain
        double qlo = qstartFit, qhi = qendFit;
        const double qmaxDom = dq * (double)(nq - 1);
        if (qhi <= qlo) { qlo = 0.0; qhi = qmaxDom; }
        long long imin, imax;
        double cppv = DEVNAN;
        if (get_window_samples(0.0, dq, nq, qlo, qhi, &imin, &imax) && imax - imin + 1 >= 2) {
            const int npts = (int)(imax - imin + 1);
            const double* yy = y + (imin - 1);                                   // yy[i], x_i = (imin - 1 + i) * dq
            double slope, intercept;
            if (npts == 2) {
                slope = (yy[1] - yy[0]) / dq;
                intercept = yy[0] - slope * ((double)(imin - 1) * dq);
            } else {
                const int numberOfPairs = npts / 2;
                const int n2 = (npts % 2 == 1) ? numberOfPairs + 1 : numberOfPairs;
                for (int i = threadIdx.x; i < numberOfPairs; i += blockDim.x) {
                    double xa = (double)(imin - 1 + i) * dq, xb = (double)(imin - 1 + n2 + i) * dq;
                    work[i] = (yy[n2 + i] - yy[i]) / (xb - xa);
                }
                {   // median of the pair slopes: sort (padded to a power of two with +inf) and apply NUMquantile (0.5)
                    int np2 = 1;
                    while (np2 < numberOfPairs) np2 <<= 1;
                    for (int i = numberOfPairs + threadIdx.x; i < np2; i += blockDim.x) work[i] = CUDART_INF;
                    __syncthreads();
                    block_bitonic_sort(work, np2);
                    slope = quantile_sorted(work, numberOfPairs, 0.5);
                    __syncthreads();
                }
                if (((npts - 1) & (npts - 2)) == 0 && npts > 2) {
                    // npts = 2^k + 1 (513 for the 1024-point cepstrum): sort the first 2^k residuals only and place the
                    // last one by comparison -- the median (element 2^(k-1)+1 of the sorted 2^k+1) is a[h-1], e or a[h]
                    const int h2 = npts - 1, h = h2 >> 1;
                    for (int i = threadIdx.x; i < h2; i += blockDim.x) work[i] = yy[i] - slope * ((double)(imin - 1 + i) * dq);
                    __syncthreads();
                    block_bitonic_sort(work, h2);
                    const double e = yy[h2] - slope * ((double)(imin - 1 + h2) * dq);
                    intercept = e <= work[h - 1] ? work[h - 1] : (e >= work[h] ? work[h] : e);
                    __syncthreads();
                } else {
                    int np2 = 1;
                    while (np2 < npts) np2 <<= 1;
                    for (int i = threadIdx.x; i < np2; i += blockDim.x)
                        work[i] = i < npts ? yy[i] - slope * ((double)(imin - 1 + i) * dq) : CUDART_INF;
                    __syncthreads();
                    block_bitonic_sort(work, np2);
                    intercept = quantile_sorted(work, npts, 0.5);
                    __syncthreads();
                }
            }
            // PowerCepstrum_getMaximumAndQuefrency: Vector_getMaximumAndX (1/ceiling, 1/floor, parabolic)
            {
                long long pmin, pmax;
                const double xlo = peakLo, xhi = peakHi;
                const bool have = get_window_samples(0.0, dq, nq, xlo, xhi, &pmin, &pmax) != 0;
                if (!have) {
                    if (threadIdx.x == 0) {
                        // no sample centre inside: linear values at the two ends
                        double il = xlo / dq + 1.0, ir = xhi / dq + 1.0;
                        int l0 = (int)floor(il), r0 = (int)floor(ir);
                        double yl = (l0 >= 1 && l0 < nq) ? y[l0 - 1] + (il - l0) * (y[l0] - y[l0 - 1]) : y[nq - 1];
                        double yr = (r0 >= 1 && r0 < nq) ? y[r0 - 1] + (ir - r0) * (y[r0] - y[r0 - 1]) : y[nq - 1];
                        s_peak[0] = yl > yr ? yl : yr;
                        s_peak[1] = yl == yr ? (xlo + xhi) / 2 : yl > yr ? xlo : xhi;
                    }
                } else {
                    // candidates in Praat's visiting order: y[pmin], y[pmax], then the parabolically refined local maxima
                    const long long lo = pmin == 1 ? 2 : pmin, hi = pmax == nq ? nq - 1 : pmax;
                    double bv = -CUDART_INF, bx = 0.0;
                    int bo = 0x7fffffff;
                    for (long long i = lo + threadIdx.x; i <= hi; i += blockDim.x) {
                        double yi = y[i - 1], ym = y[i - 2], yp = y[i];
                        if (yi > ym && yi >= yp) {
                            double dy = 0.5 * (yp - ym), d2y = 2 * yi - ym - yp;
                            double loc = yi + 0.5 * dy * dy / d2y;
                            int ord = 2 + (int)(i - lo);
                            if (loc > bv || (loc == bv && ord < bo)) { bv = loc; bx = (double)i + dy / d2y; bo = ord; }
                        }
                    }
                    if (threadIdx.x == 0) {
                        double e0 = y[pmin - 1], e1 = y[pmax - 1];
                        double ev = e0, ex = (double)pmin;
                        int eo = 0;
                        if (e1 > ev) { ev = e1; ex = (double)pmax; eo = 1; }
                        if (ev > bv || (ev == bv && eo < bo)) { bv = ev; bx = ex; bo = eo; }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        double ov = __shfl_xor_sync(FULL_MASK, bv, o), ox = __shfl_xor_sync(FULL_MASK, bx, o);
                        int oo = __shfl_xor_sync(FULL_MASK, bo, o);
                        if (ov > bv || (ov == bv && oo < bo)) { bv = ov; bx = ox; bo = oo; }
                    }
                    if ((threadIdx.x & 31) == 0) { s_pv[threadIdx.x >> 5] = bv; s_px[threadIdx.x >> 5] = bx; s_po[threadIdx.x >> 5] = bo; }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
                            if (s_pv[w] > bv || (s_pv[w] == bv && s_po[w] < bo)) { bv = s_pv[w]; bx = s_px[w]; bo = s_po[w]; }
                        double x = (bx - 1.0) * dq;
                        if (x < xlo) x = xlo; else if (x > xhi) x = xhi;
                        s_peak[0] = bv; s_peak[1] = x;
                    }
                }
            }
            __syncthreads();
            cppv = s_peak[0] - (slope * s_peak[1] + intercept);
        }
        if (threadIdx.x == 0) cpp_frame[f] = cppv;
    }
}

// ---- warp-per-frame variant (the usual 1024-point cepstrum: 513 quefrency bins) -----------------------------------------
// The two medians of the robust line fit dominate the frame: 256 pair slopes and 513 residuals.  Here one warp owns a frame
// and sorts in REGISTERS (8 / 16 values per lane, bitonic network: exchanges at distance < E stay inside a lane, the others
// are one shuffle per value), so no block barrier is ever taken and eight frames advance independently per CTA.
template <int E>
__device__ __forceinline__ void warp_bitonic_sort(double (&v)[E], int lane) {
    constexpr int LOGN = E == 8 ? 8 : 9;
    static_assert(E == 8 || E == 16, "8 or 16 values per lane");
#pragma unroll
    for (int lk = 1; lk <= LOGN; lk++) {
        const int k = 1 << lk;
#pragma unroll
        for (int lj = lk - 1; lj >= 0; lj--) {
            const int j = 1 << lj;
            if (j >= E) {
                const int pl = j / E;
                const bool up = ((lane * E) & k) == 0;
                const bool keep_min = ((lane & pl) == 0) == up;
#pragma unroll
                for (int e = 0; e < E; e++) {
                    const double o = __shfl_xor_sync(FULL_MASK, v[e], pl);
                    v[e] = ((o < v[e]) == keep_min) ? o : v[e];          // no NaNs here; on a tie either copy will do
                }
            } else {
#pragma unroll
                for (int e = 0; e < E; e++) {
                    if ((e & j) == 0) {
                        const bool up = (((lane * E) | e) & k) == 0;
                        const double a = v[e], b = v[e | j];
                        const bool sw = (a > b) == up;
                        v[e] = sw ? b : a;
                        v[e | j] = sw ? a : b;
                    }
                }
            }
        }
    }
}

#define CPW 8                       // frames (warps) per CTA
#define CPW_STRIDE 1048             // doubles of shared memory per warp: col[524] + y[524]
__global__ void __launch_bounds__(32 * CPW, 2) k_cpp_frames_warp(const CepSeg* __restrict__ segs, const int* __restrict__ fprefix,
                                                                  int nsegs, const double* __restrict__ cep, int nqmax, int nTimeAvg,
                                                                  double qAvgWindow, double peakLo, double peakHi, double qstartFit,
                                                                  double qendFit, double* __restrict__ cpp_frame) {
    extern __shared__ __align__(16) unsigned char cpw_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* col = (double*)cpw_smem + (size_t)wid * CPW_STRIDE;       // time-averaged column, later the sorted values
    double* y = col + CPW_STRIDE / 2;
    const int total = fprefix[nsegs];
    for (int f = blockIdx.x * CPW + wid; f < total; f += gridDim.x * CPW) {
        __syncwarp();
        const int sgi = find_segment(fprefix, nsegs, f);
        const CepSeg S = segs[sgi];
        const int f0 = fprefix[sgi], nFrames = fprefix[sgi + 1] - f0;
        const int i1 = f - f0 + 1;                                               // 1-based frame
        const int nq = S.nfft / 2 + 1;
        const double dq = S.dq;
        int jfrom = i1, jto = i1;
        if (nTimeAvg > 1) {
            jfrom = i1 - nTimeAvg / 2; jto = i1 + nTimeAvg / 2;
            if ((nTimeAvg % 2) == 0) jto--;
            if (jfrom < 1) jfrom = 1;
            if (jto > nFrames) jto = nFrames;
        }
        if (jto - jfrom < 5) {
            // up to five rows (the 0.01 s window at 2 ms steps): all loads of a bin are issued before the first addition;
            // a row beyond jto contributes an exact 0.0 to the same left-to-right sum
            const double* r0 = cep + (size_t)(f0 + jfrom - 1) * nqmax;
            const int nrow = jto - jfrom + 1;
#pragma unroll 2
            for (int iq = lane; iq < nq; iq += 32) {
                double r[5];
#pragma unroll
                for (int t = 0; t < 5; t++) r[t] = t < nrow ? __ldg(r0 + (size_t)t * nqmax + iq) : 0.0;
                double s = 0.0;
#pragma unroll
                for (int t = 0; t < 5; t++) s += r[t];
                col[iq] = nTimeAvg > 1 ? s / (double)nrow : s;
            }
        } else {
            for (int iq = lane; iq < nq; iq += 32) {
                double s = 0.0;
                for (int j = jfrom; j <= jto; j++) s += __ldg(cep + (size_t)(f0 + j - 1) * nqmax + iq);
                col[iq] = nTimeAvg > 1 ? s / (double)(jto - jfrom + 1) : s;
            }
        }
        __syncwarp();
        const int nQ = (int)floor(qAvgWindow / dq);
        for (int iq = lane; iq < nq; iq += 32) {
            double v;
            if (nQ > 1) {
                int i = iq + 1;
                int qf = i - nQ / 2, qt = i + nQ / 2;
                if ((nQ % 2) == 0) qt--;
                if (qf < 1) qf = 1;
                if (qt > nq) qt = nq;
                double s = 0.0;
                for (int j = qf; j <= qt; j++) s += col[j - 1];
                v = s / (double)(qt - qf + 1);
            } else v = col[iq];
            y[iq] = 10.0 * log10(v + 1e-30);
        }
        __syncwarp();
        double qlo = qstartFit, qhi = qendFit;
        const double qmaxDom = dq * (double)(nq - 1);
        if (qhi <= qlo) { qlo = 0.0; qhi = qmaxDom; }
        long long imin, imax;
        double cppv = DEVNAN;
        if (get_window_samples(0.0, dq, nq, qlo, qhi, &imin, &imax) && imax - imin + 1 >= 2) {
            const int npts = (int)(imax - imin + 1);
            const double* yy = y + (imin - 1);
            double slope, intercept;
            if (npts == 2) {
                slope = (yy[1] - yy[0]) / dq;
                intercept = yy[0] - slope * ((double)(imin - 1) * dq);
            } else {
                const int numberOfPairs = npts / 2;
                const int n2 = (npts % 2 == 1) ? numberOfPairs + 1 : numberOfPairs;
                {   // median of the pair slopes (the order in which the values enter the network is irrelevant)
                    double v[8];
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const int i = e * 32 + lane;
                        double xa = (double)(imin - 1 + i) * dq, xb = (double)(imin - 1 + n2 + i) * dq;
                        v[e] = i < numberOfPairs ? (yy[n2 + i] - yy[i]) / (xb - xa) : CUDART_INF;
                    }
                    warp_bitonic_sort<8>(v, lane);
                    if (numberOfPairs == 256) {
                        const double a127 = __shfl_sync(FULL_MASK, v[7], 15), a128 = __shfl_sync(FULL_MASK, v[0], 16);
                        slope = a128 == a127 ? a127 : a127 + 0.5 * (a128 - a127);     // NUMquantile (0.5): place 128.5
                    } else {
                        __syncwarp();
#pragma unroll
                        for (int e = 0; e < 8; e++) col[lane * 8 + e] = v[e];
                        __syncwarp();
                        slope = quantile_sorted(col, numberOfPairs, 0.5);
                    }
                }
                {   // median of the residuals
                    const bool pow2p1 = npts == 513;
                    const int nsort = pow2p1 ? 512 : npts;
                    double v[16];
#pragma unroll
                    for (int e = 0; e < 16; e++) {
                        const int i = e * 32 + lane;
                        v[e] = i < nsort ? yy[i] - slope * ((double)(imin - 1 + i) * dq) : CUDART_INF;
                    }
                    warp_bitonic_sort<16>(v, lane);
                    if (pow2p1) {
                        // 513 values: sort the first 512 and place the last one by comparison -- the median is a[255], e or a[256]
                        const double lo = __shfl_sync(FULL_MASK, v[15], 15), hi = __shfl_sync(FULL_MASK, v[0], 16);
                        const double e = yy[512] - slope * ((double)(imin - 1 + 512) * dq);
                        intercept = e <= lo ? lo : (e >= hi ? hi : e);
                    } else {
                        __syncwarp();
#pragma unroll
                        for (int e = 0; e < 16; e++) col[lane * 16 + e] = v[e];
                        __syncwarp();
                        intercept = quantile_sorted(col, npts, 0.5);
                    }
                }
            }
            // PowerCepstrum_getMaximumAndQuefrency: Vector_getMaximumAndX (1/ceiling, 1/floor, parabolic)
            double pk_v, pk_x;
            {
                long long pmin, pmax;
                const double xlo = peakLo, xhi = peakHi;
                if (!get_window_samples(0.0, dq, nq, xlo, xhi, &pmin, &pmax)) {
                    double il = xlo / dq + 1.0, ir = xhi / dq + 1.0;
                    int l0 = (int)floor(il), r0 = (int)floor(ir);
                    double yl = (l0 >= 1 && l0 < nq) ? y[l0 - 1] + (il - l0) * (y[l0] - y[l0 - 1]) : y[nq - 1];
                    double yr = (r0 >= 1 && r0 < nq) ? y[r0 - 1] + (ir - r0) * (y[r0] - y[r0 - 1]) : y[nq - 1];
                    pk_v = yl > yr ? yl : yr;
                    pk_x = yl == yr ? (xlo + xhi) / 2 : yl > yr ? xlo : xhi;
                } else {
                    const long long lo = pmin == 1 ? 2 : pmin, hi = pmax == nq ? nq - 1 : pmax;
                    double bv = -CUDART_INF, bx = 0.0;
                    int bo = 0x7fffffff;
                    for (long long i = lo + lane; i <= hi; i += 32) {
                        double yi = y[i - 1], ym = y[i - 2], yp = y[i];
                        if (yi > ym && yi >= yp) {
                            double dy = 0.5 * (yp - ym), d2y = 2 * yi - ym - yp;
                            double loc = yi + 0.5 * dy * dy / d2y;
                            int ord = 2 + (int)(i - lo);
                            if (loc > bv || (loc == bv && ord < bo)) { bv = loc; bx = (double)i + dy / d2y; bo = ord; }
                        }
                    }
                    if (lane == 0) {
                        double e0 = y[pmin - 1], e1 = y[pmax - 1];
                        double ev = e0, ex = (double)pmin;
                        int eo = 0;
                        if (e1 > ev) { ev = e1; ex = (double)pmax; eo = 1; }
                        if (ev > bv || (ev == bv && eo < bo)) { bv = ev; bx = ex; bo = eo; }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        double ov = __shfl_xor_sync(FULL_MASK, bv, o), ox = __shfl_xor_sync(FULL_MASK, bx, o);
                        int oo = __shfl_xor_sync(FULL_MASK, bo, o);
                        if (ov > bv || (ov == bv && oo < bo)) { bv = ov; bx = ox; bo = oo; }
                    }
                    double x = (bx - 1.0) * dq;
                    if (x < xlo) x = xlo; else if (x > xhi) x = xhi;
                    pk_v = bv; pk_x = x;
                }
            }
            cppv = pk_v - (slope * pk_x + intercept);
        }
        if (lane == 0) cpp_frame[f] = cppv;
    }
}

// CPPS per segment = mean over frames; clip value = mean of the segments with CPPS > 4 (mshds_extractor.py:293,298)
__global__ void __launch_bounds__(128) k_cpp_reduce(Clips c, CppSegs sg, const int* __restrict__ seg_prefix,
                                                     const int* __restrict__ fprefix, const double* __restrict__ cpp_frame) {
    __shared__ double red[32];
    const int clip = blockIdx.x;
    const int s0 = seg_prefix[clip], ns = seg_prefix[clip + 1] - s0;
    double sum = 0.0;
    int cnt = 0;
    for (int k = 0; k < ns; k++) {
        const int f0 = fprefix[s0 + k], nF = fprefix[s0 + k + 1] - f0;
        double a = 0.0;
        bool bad = false;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) { double v = cpp_frame[f0 + i]; if (is_undef(v)) bad = true; else a += v; }
        a = block_sum(a, red);
        double nb = block_sum(bad ? 1.0 : 0.0, red);
        if (nF >= 1 && nb == 0.0) {
            double cpps = a / (double)nF;
            if (cpps > 4.0) { sum += cpps; cnt++; }
        }
    }
    if (threadIdx.x == 0) {
        const bool fail = sg.fail[clip] != 0;
        c.feat[(size_t)clip * N_FEAT + 12] = (!fail && cnt > 0) ? sum / (double)cnt : DEVNAN;
        if (fail || cnt == 0) atomicOr(&c.status[clip], ST_CPP);
    }
}

void launch_vuv_segments(const Clips& c, const PulseSet& ps, const CppSegs& sg, cudaStream_t s) {
    k_vuv_segments<<<(c.n + 63) / 64, 64, 0, s>>>(c, ps, sg, 0.02, 0.1);
}
void launch_cepstrogram(const CepSeg* segs, const int* fprefix, int nsegs, const ResampleJob* jobs, const double* sig,
                        const double2* tw, double emphasis, double dt, double* cep, int nqmax, int total_frames,
                        const double* wtab, int wtab_n, cudaStream_t s, const double2* twb512, int* turn_counter) {
    int skip = 0;
    if (twb512 && turn_counter && nqmax >= 513) {
        // frames of the usual 1024-point transform: warp-per-frame register FFT; the rest (short segments) below
        cudaMemsetAsync(turn_counter, 0, sizeof(int), s);
        const size_t smem_w = (size_t)CEW_TW_BYTES + (size_t)CEW_WARPS * 1024 * 16;
        cudaFuncSetAttribute(k_cepstrogram_w, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w);
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cepstrogram_w, 32 * CEW_WARPS, smem_w);
        if (occ < 1) occ = 1;
        int gw = sm_count() * occ;
        const int nt2 = (total_frames + 2 * CEW_WARPS - 1) / (2 * CEW_WARPS);
        if (gw > nt2) gw = nt2;
        if (gw < 1) gw = 1;
        k_cepstrogram_w<<<gw, 32 * CEW_WARPS, smem_w, s>>>(segs, fprefix, nsegs, jobs, sig, tw, twb512, emphasis, dt, cep, nqmax, wtab, wtab_n,
                                                         turn_counter);
        skip = 1024;
    }
    static int nt = 0;
    if (!nt) { const char* e = getenv("MSHDS_NT_CEP"); nt = e && atoi(e) == 256 ? 256 : (e && atoi(e) == 128 ? 128 : NT_CEP_DEFAULT); }   // development switch
    const int cap = nt == 128 ? sm_count() * 12 : sm_count() * 8;
    int grid = total_frames < cap ? total_frames : cap;
    if (grid < 1) grid = 1;
    size_t smem = sizeof(double2) * 512 + sizeof(double) * 32;
    k_cepstrogram<<<grid, nt, smem, s>>>(segs, fprefix, nsegs, jobs, sig, tw, emphasis, dt, cep, nqmax, wtab, wtab_n, skip);
}
void launch_cpp_frames(const CepSeg* segs, const int* fprefix, int nsegs, const double* cep, int nqmax, int nTimeAvg,
                       double qAvgWindow, double* cpp_frame, int total_frames, cudaStream_t s) {
    int grid = total_frames < sm_count() * 8 ? total_frames : sm_count() * 8;
    if (grid < 1) grid = 1;
    static int use_block = -1;
    if (use_block < 0) { const char* e = getenv("MSHDS_CPP_BLOCK"); use_block = e && atoi(e) ? 1 : 0; }    // development switch
    if (nqmax <= 513 && !use_block) {
        const size_t smem = sizeof(double) * CPW_STRIDE * CPW;
        int g = (total_frames + CPW - 1) / CPW;
        if (g > sm_count() * 2) g = sm_count() * 2;
        if (g < 1) g = 1;
        cudaFuncSetAttribute(k_cpp_frames_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_cpp_frames_warp<<<g, 32 * CPW, smem, s>>>(segs, fprefix, nsegs, cep, nqmax, nTimeAvg, qAvgWindow, 1.0 / 330.0, 1.0 / 60.0,
                                                   0.001, 0.0, cpp_frame);
        return;
    }
    k_cpp_frames<<<grid, 256, 0, s>>>(segs, fprefix, nsegs, cep, nqmax, nTimeAvg, qAvgWindow, 1.0 / 330.0, 1.0 / 60.0, 0.001, 0.0,
                                      cpp_frame);
}
void launch_cpp_reduce(const Clips& c, const CppSegs& sg, const int* seg_prefix, const int* fprefix, const double* cpp_frame,
                       cudaStream_t s) {
    k_cpp_reduce<<<c.n, 128, 0, s>>>(c, sg, seg_prefix, fprefix, cpp_frame);
}
