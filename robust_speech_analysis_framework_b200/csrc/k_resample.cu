// k_resample.cu -- Sound_resample (fon/Sound.cpp): FFT brick-wall low-pass of the zero-padded sound followed by
// windowed-sinc interpolation at the new sampling times.
//
// Serves the resample-to-10-kHz inside "To Formant (burg)" (precision 500, whole clip, mshds_extractor.py:319) and
// inside "To PowerCepstrogram" (precision 50, one voiced segment at a time, :289).
//
// The low-pass needs one FFT of nfft = 2^k >= nx + 2000 points per sound (2^19 for a 30 s clip, 2^24 for 10 min).  It is
// a hand-written multi-pass FFT: strided radix-2^b passes (<= 256 points per column, 16 columns per CTA so every global
// access is a 256-byte row segment) down to contiguous 4096-point blocks; the innermost kernel runs forward DIF ->
// Praat's packed-spectrum zeroing -> inverse DIT on the block without leaving shared memory; the strided passes are then
// undone in reverse.  Forward output is bit-reversed and the inverse consumes bit-reversed input, so no transposition
// pass exists.  The first pass converts the int16 samples on load, the last one writes the filtered real samples.
#include "internal.h"
#include "common.cuh"
#include "fft.cuh"

#define RS_NTHR 256
#define TB 16                      // columns per CTA in the strided passes
#define ANTI_TURN 1000             // Praat's antiTurnAround

__device__ __forceinline__ double job_src(const ResampleJob& J, const SPtr& pcm, long long pos /*0-based in data*/) {
    long long i = pos - (ANTI_TURN - 1);                 // data[antiTurnAround + i] = z[i], i = 1..nx
    if (i < 1 || i > J.nx) return 0.0;
    long long ic = J.ix1 + i - 1;                        // sample index within the clip (1-based); outside -> virtual zero
    if (ic < 1 || ic > J.nclip) return 0.0;
    return samp(pcm + J.clip_off, ic - 1);
}

// One strided DIF / DIT pass over blocks of length Nl = 2^nl: R = 2^rb point transforms at stride M = Nl / R.
template <bool INVERSE, bool FIRST_FROM_SRC, bool LAST_TO_REAL>
__global__ void __launch_bounds__(RS_NTHR) k_fft_strided(const ResampleJob* __restrict__ jobs, const int* __restrict__ ids,
                                                          SPtr pcm, double2* __restrict__ zbuf,
                                                          double* __restrict__ filt, const double2* __restrict__ tw,
                                                          int logn, int nl, int rb, int mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* a = (double2*)smem;
    const int R = 1 << rb;
    const long long N = 1LL << logn, Nl = 1LL << nl, M = Nl >> rb;
    const long long tilesPerJob = N / ((long long)R * TB);
    const long long tile = blockIdx.x % tilesPerJob;
    const ResampleJob J = jobs[ids[blockIdx.x / tilesPerJob]];
    const long long tilesPerBlk = M / TB;
    const long long blk = tile / tilesPerBlk, col0 = (tile % tilesPerBlk) * TB;
    double2* z = zbuf + J.zoff + blk * Nl;
    for (int e = threadIdx.x; e < R * TB; e += RS_NTHR) {
        int r = e / TB, b = e % TB;
        long long pos = (long long)r * M + col0 + b;
        double2 v;
        if (FIRST_FROM_SRC) v = make_double2(job_src(J, pcm, blk * Nl + pos), 0.0);
        else v = z[pos];
        if (INVERSE) {
            int cidx = bitrev(r, rb);
            if (cidx != 0) {
                double s, co;
                sincospi(2.0 * (double)((col0 + b) * cidx) / (double)Nl, &s, &co);
                v = cmul(v, make_double2(co, s));
            }
        }
        a[e] = v;
    }
    __syncthreads();
    if (INVERSE) fft_dit_cols<+1>(a, R, TB, tw);
    else fft_dif_cols<-1>(a, R, TB, tw);
    for (int e = threadIdx.x; e < R * TB; e += RS_NTHR) {
        int r = e / TB, b = e % TB;
        long long pos = (long long)r * M + col0 + b;
        double2 v = a[e];
        if (!INVERSE) {
            int cidx = bitrev(r, rb);
            if (cidx != 0) {
                double s, co;
                sincospi(-2.0 * (double)((col0 + b) * cidx) / (double)Nl, &s, &co);
                v = cmul(v, make_double2(co, s));
            }
            z[pos] = v;
        } else if (LAST_TO_REAL) {
            long long gp = blk * Nl + pos;
            long long i = gp - ANTI_TURN;                    // to[i] = data[i + antiTurnAround] / nfft, i = 1..nx (0-based i here)
            if (i >= 0 && i < J.nx) {
                if (mode == 0) filt[J.filt_off + i] = v.x * (1.0 / (double)N);
                else filt[J.out_off + 2 * i + (mode - 1)] = v.x * (1.0 / (double)N);      // `filt` is the output sound here
            }
        } else {
            z[pos] = v;
        }
    }
}

// Praat: "for (i = floor(upfactor*nfft); i <= nfft; i++) data[i] = 0;  data[2] = 0;" on the NUMrealft-packed spectrum
// (data[1] = DC, data[2] = Nyquist, data[2k+1] = Re X_k, data[2k+2] = Im X_k).  Returns the masked bin.
// mode 1 / 2 = Sound_upsample (exact doubling): packed values data[i], i > imin = (long)(0.95 N), are tapered by
// (N - i) / (N - imin), data[2] (Nyquist) = 0, and the inverse transform has twice the length.  Its even output samples are
// the length-N inverse of the tapered spectrum (mode 1), the odd ones the inverse of the spectrum advanced by half a sample
// (mode 2: bin k times exp(+i pi k / N), k counted as a signed frequency).
__device__ __forceinline__ double2 apply_lowpass_mask(double2 v, long long k, long long N, long long i0, int mode) {
    long long kk = k <= N / 2 ? k : N - k;
    if (mode != 0) {
        if (kk == 0) return v;
        if (kk == N / 2) return make_double2(0.0, 0.0);
        const bool neg = k > N / 2;
        double2 X = neg ? make_double2(v.x, -v.y) : v;                 // the bin of the positive frequency kk
        const long long imin = (long long)((double)N * 0.95);
        const long long ir = 2 * kk + 1, ii = 2 * kk + 2;
        if (ir > imin) X.x *= (double)(N - ir) / (double)(N - imin);
        if (ii > imin) X.y *= (double)(N - ii) / (double)(N - imin);
        if (mode == 2) {
            double sn, cs;
            sincospi((double)kk / (double)N, &sn, &cs);
            X = cmul(X, make_double2(cs, sn));
        }
        return neg ? make_double2(X.x, -X.y) : X;
    }
    if (kk == 0) { if (i0 <= 1) v.x = 0.0; v.y = v.y; return v; }
    if (kk == N / 2) return make_double2(0.0, 0.0);
    if (2 * kk + 1 >= i0) v.x = 0.0;
    if (2 * kk + 2 >= i0) v.y = 0.0;
    return v;
}

// Innermost kernel: contiguous blocks of Cn = 2^cb points: forward DIF, mask, inverse DIT, all in shared memory.
template <bool WHOLE>     // WHOLE: logn == cb (load from the source, write filtered reals)
__global__ void __launch_bounds__(RS_NTHR) k_fft_inner(const ResampleJob* __restrict__ jobs, const int* __restrict__ ids,
                                                        SPtr pcm, double2* __restrict__ zbuf,
                                                        double* __restrict__ filt, const double2* __restrict__ tw, int logn,
                                                        int cb, double upfactor, int mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* a = (double2*)smem;
    const int Cn = 1 << cb;
    const long long N = 1LL << logn;
    const long long chunksPerJob = N >> cb;
    const long long chunk = blockIdx.x % chunksPerJob;
    const ResampleJob J = jobs[ids[blockIdx.x / chunksPerJob]];
    double2* z = zbuf + J.zoff + chunk * Cn;
    for (int e = threadIdx.x; e < Cn; e += RS_NTHR) a[SWZ(e)] = WHOLE ? make_double2(job_src(J, pcm, e), 0.0) : z[e];
    __syncthreads();
    fft_dif<-1>(a, Cn, tw);
    const long long i0 = (long long)floor(upfactor * (double)N);
    for (int e = threadIdx.x; e < Cn; e += RS_NTHR) {
        long long pos = chunk * Cn + e;
        long long k = (long long)(__brevll((unsigned long long)pos) >> (64 - logn));
        a[SWZ(e)] = apply_lowpass_mask(a[SWZ(e)], k, N, i0, mode);
    }
    __syncthreads();
    fft_dit<+1>(a, Cn, tw);
    for (int e = threadIdx.x; e < Cn; e += RS_NTHR) {
        if (WHOLE) {
            long long i = (long long)e - ANTI_TURN;
            if (i >= 0 && i < J.nx) {
                if (mode == 0) filt[J.filt_off + i] = a[SWZ(e)].x * (1.0 / (double)N);
                else filt[J.out_off + 2 * i + (mode - 1)] = a[SWZ(e)].x * (1.0 / (double)N);
            }
        } else z[e] = a[SWZ(e)];
    }
}

// Up-sampling (upfactor >= 1) skips the low-pass: the interpolation reads the source samples themselves.
__global__ void __launch_bounds__(256) k_copy_src(const ResampleJob* __restrict__ jobs, int njobs, SPtr pcm, double* __restrict__ filt) {
    for (int jb = blockIdx.y; jb < njobs; jb += gridDim.y) {
        const ResampleJob J = jobs[jb];
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.nx; i += (long long)gridDim.x * blockDim.x)
            filt[J.filt_off + i] = job_src(J, pcm, i + ANTI_TURN);
    }
}
void launch_resample_copy(const ResampleJob* d_jobs, int njobs, long long max_nx, SPtr pcm, double* filt, cudaStream_t s, long long* launches) {
    if (njobs < 1) return;
    long long bx = (max_nx + 255) / 256;
    if (bx > 4096) bx = 4096;
    if (bx < 1) bx = 1;
    k_copy_src<<<dim3((unsigned)bx, (unsigned)(njobs < 65535 ? njobs : 65535)), 256, 0, s>>>(d_jobs, njobs, pcm, filt);
    (*launches)++;
}

// ------------------------------------------------------------------------------------------------ sinc interpolation
// Coefficient table of NUM_interpolate_sinc for the P distinct fractional positions of a rational rate change.
__global__ void k_sinc_table(const ResampleJob* __restrict__ jobs, const int* __restrict__ rep, int ntables,
                             double* __restrict__ table, double* __restrict__ table_fl, long long* __restrict__ table_mid,
                             int P, int D, double dx_src) {
    const int tab = blockIdx.x;
    if (tab >= ntables) return;
    const ResampleJob J = jobs[rep[tab]];
    double* T = table + (size_t)tab * P * 2 * D;
    for (int ph = threadIdx.x; ph < P; ph += blockDim.x) {
        long long j = (long long)(J.nout / 2 / P) * P + ph + 1;
        double x = J.out_x1 + (double)(j - 1) * J.out_dx;
        double index = (x - J.x1) / dx_src + 1.0;
        table_fl[(size_t)tab * P + ph] = index - floor(index);
        table_mid[(size_t)tab * P + ph] = (long long)floor(index);
    }
    for (int e = threadIdx.x; e < P * 2 * D; e += blockDim.x) {
        int ph = e / (2 * D), k = e % (2 * D);
        // representative output sample of this phase near the middle of the sound
        long long j = (long long)(J.nout / 2 / P) * P + ph + 1;
        double x = J.out_x1 + (double)(j - 1) * J.out_dx;
        double index = (x - J.x1) / dx_src + 1.0;
        double fl = index - floor(index), fr = 1.0 - fl;
        double v;
        if (fl == 0.0) v = (k == 0) ? 1.0 : 0.0;
        else if (k < D) {          // left tap k: sample midleft - k
            double sgn = (k & 1) ? -1.0 : 1.0;
            double al = fl + k;
            v = sgn * (0.5 * sinpi(fl)) / (MSHDS_PI * al) * (1.0 + cospi(al * (1.0 / (fl + D))));
        } else {                   // right tap kk: sample midright + kk
            int kk = k - D;
            double sgn = (kk & 1) ? -1.0 : 1.0;
            double ar = fr + kk;
            v = sgn * (0.5 * sinpi(fr)) / (MSHDS_PI * ar) * (1.0 + cospi(ar * (1.0 / (fr + D))));
        }
        T[e] = v;
    }
}

// exact NUM_interpolate_sinc, one warp per evaluation (edge samples where the depth is clipped, and samples whose
// fractional position falls on the other side of an integer than the table's representative)
__device__ double sinc_interp_warp_ll(const double* __restrict__ y /*1-based*/, long long n, double x, long long maxDepth, int lane) {
    long long midleft = (long long)floor(x), midright = midleft + 1;
    if (n < 1) return DEVNAN;
    if (x > n) return y[n];
    if (x < 1) return y[1];
    if (x == midleft) return y[midleft];
    if (maxDepth > midright - 1) maxDepth = midright - 1;
    if (maxDepth > n - midleft) maxDepth = n - midleft;
    if (maxDepth <= 0) return y[(long long)floor(x + 0.5)];
    if (maxDepth == 1) return y[midleft] + (x - midleft) * (y[midright] - y[midleft]);
    if (maxDepth == 2) {
        double yl = y[midleft], yr = y[midright];
        double dyl = 0.5 * (yr - y[midleft - 1]), dyr = 0.5 * (y[midright + 1] - yl);
        double fil = x - midleft, fir = midright - x;
        return yl * fir + yr * fil - fil * fir * (0.5 * (dyr - dyl) + (fil - 0.5) * (dyl + dyr - 2 * (yr - yl)));
    }
    double fl = x - midleft, fr = midright - x;
    double hsl = 0.5 * sinpi(fl), hsr = 0.5 * sinpi(fr);
    double invl = 1.0 / (fl + maxDepth), invr = 1.0 / (fr + maxDepth);
    double acc = 0.0;
    for (long long k = lane; k < maxDepth; k += 32) {
        double sgn = (k & 1) ? -1.0 : 1.0;
        double al = fl + k, ar = fr + k;
        acc += y[midleft - k] * (sgn * hsl / (MSHDS_PI * al) * (1.0 + cospi(al * invl)));
        acc += y[midright + k] * (sgn * hsr / (MSHDS_PI * ar) * (1.0 + cospi(ar * invr)));
    }
    return warp_sum(acc);
}

// One warp per output sample: lanes split the 2*D taps (coalesced reads of the filtered signal and the table row).
__global__ void __launch_bounds__(256) k_sinc_apply(const ResampleJob* __restrict__ jobs, const long long* __restrict__ out_prefix,
                                                     int njobs, const double* __restrict__ filt, const double* __restrict__ table,
                                                     const double* __restrict__ table_fl, double* __restrict__ out, int P, int D,
                                                     double dx_src) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total = out_prefix[njobs];
    for (long long o = gw; o < total; o += nw) {
        const int job = find_segment_ll(out_prefix, njobs, o);
        const ResampleJob J = jobs[job];
        const long long j = o - out_prefix[job] + 1;                   // 1-based output sample
        const double x = J.out_x1 + (double)(j - 1) * J.out_dx;
        const double index = (x - J.x1) / dx_src + 1.0;
        const double* y = filt + J.filt_off - 1;                        // 1-based
        const long long midleft = (long long)floor(index);
        double v;
        bool fast = P > 0 && midleft >= D && midleft + D <= J.nx && index != (double)midleft;
        if (fast) {
            // the table row was built for one representative fractional position; a sample whose floating-point index
            // lands on the other side of an integer (fraction ~0 vs ~1) must take the exact path
            double tfl = table_fl[(size_t)J.table_id * P + (size_t)((j - 1) % P)];
            fast = fabs((index - (double)midleft) - tfl) < 1e-8;
        }
        if (fast) {
            const double* T = table + ((size_t)J.table_id * P + (size_t)((j - 1) % P)) * 2 * D;
            double acc = 0.0;
            for (int k = lane; k < D; k += 32) {
                acc = fma(y[midleft - k], T[k], acc);
                acc = fma(y[midleft + 1 + k], T[D + k], acc);
            }
            v = warp_sum(acc);
        } else {
            v = sinc_interp_warp_ll(y, J.nx, index, D, lane);
        }
        if (lane == 0) out[J.out_off + j - 1] = v;
    }
}

// ------------------------------------------------------------------------------------------------ polyphase FIR
// Regular outputs of a rational rate change: output j = P*m + ph + 1 has midleft = mid_rep(ph) + Q*(m - m_rep) and one of P
// fixed coefficient rows, i.e. P FIR filters over Q-decimated input streams.  One warp per phase, one thread computes FIR_R
// consecutive outputs of its phase with a sliding register window: per tap step 2 shared-memory loads feed FIR_R FMAs.
// The input tile is staged de-interleaved (Q streams) so lane strides are FIR_R doubles (odd => conflict-free).
#define FIR_R 7
__global__ void __launch_bounds__(512) k_sinc_fir(const ResampleJob* __restrict__ jobs, const int* __restrict__ tile_prefix, int njobs,
                                                   const double* __restrict__ filt, const double* __restrict__ table,
                                                   const long long* __restrict__ table_mid, double* __restrict__ out, int P, int Q, int D) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* U = (double*)smem_raw;
    __shared__ int s_job;
    if (threadIdx.x == 0) s_job = find_segment(tile_prefix, njobs, blockIdx.x);
    __syncthreads();
    const int job = s_job;
    const ResampleJob J = jobs[job];
    const int tt = blockIdx.x - tile_prefix[job];
    const long long m_rep = J.nout / 2 / P;
    const long long m0 = (long long)tt * (32 * FIR_R);
    const long long* mids = table_mid + (size_t)J.table_id * P;
    // tile sample range [lo, hi] (1-based indices into the filtered signal)
    long long mid_min = 0x7fffffffffffffffLL, mid_max = -0x7fffffffffffffffLL;
    for (int ph = 0; ph < P; ph++) {
        long long mm = mids[ph] + (long long)Q * (m0 - m_rep);
        mid_min = mm < mid_min ? mm : mid_min;
        mid_max = mm > mid_max ? mm : mid_max;
    }
    const long long lo = mid_min - (D - 1) - Q;                     // one extra Q of head-room for the sliding window
    const long long hi = mid_max + (long long)Q * (32 * FIR_R - 1) + D + Q;
    const int span = (int)(hi - lo + 1);
    const int NU = ((span + Q - 1) / Q + 2) | 1;                     // odd row length
    const double* y = filt + J.filt_off - 1;                         // 1-based
    for (int pos = threadIdx.x; pos < span; pos += blockDim.x) {
        long long i = lo + pos;
        U[(pos % Q) * NU + pos / Q] = (i >= 1 && i <= J.nx) ? y[i] : 0.0;
    }
    __syncthreads();
    const int ph = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (ph >= P) return;
    const double* T = table + ((size_t)J.table_id * P + ph) * 2 * D;
    const long long mid0 = mids[ph] + (long long)Q * (m0 - m_rep);   // midleft of output m0 of this phase
    const int off = (int)(mid0 - lo);                                // position of that sample in the tile
    double acc[FIR_R];
#pragma unroll
    for (int r = 0; r < FIR_R; r++) acc[r] = 0.0;
    // left taps: sample mid - k, k = k0 + Q*s
    for (int k0 = 0; k0 < Q && k0 < D; k0++) {
        const int e = off - k0;
        const double* Us = U + (e % Q) * NU + e / Q + lane * FIR_R;
        double w[FIR_R];
#pragma unroll
        for (int r = 0; r < FIR_R; r++) w[r] = Us[r];
        // the window is a ring (sample Us[i] in slot i mod FIR_R): FIR_R tap steps per trip, no register shuffling
        int sidx = 0, k = k0;
        for (; k + (FIR_R - 1) * Q < D; k += FIR_R * Q) {
#pragma unroll
            for (int t = 0; t < FIR_R; t++) {
                const double cf = __ldg(T + k + t * Q);
#pragma unroll
                for (int r = 0; r < FIR_R; r++) acc[r] = fma(cf, w[(r - t + FIR_R) % FIR_R], acc[r]);
                sidx++;
                w[(FIR_R - 1 - t) % FIR_R] = Us[-sidx];
            }
        }
        for (; k < D; k += Q) {          // tail (< FIR_R steps): slots are aligned again, shift the window
            const double cf = __ldg(T + k);
#pragma unroll
            for (int r = 0; r < FIR_R; r++) acc[r] = fma(cf, w[r], acc[r]);
#pragma unroll
            for (int r = FIR_R - 1; r > 0; r--) w[r] = w[r - 1];
            sidx++;
            w[0] = Us[-sidx];
        }
    }
    // right taps: sample mid + 1 + k
    for (int k0 = 0; k0 < Q && k0 < D; k0++) {
        const int e = off + 1 + k0;
        const double* Us = U + (e % Q) * NU + e / Q + lane * FIR_R;
        double w[FIR_R];
#pragma unroll
        for (int r = 0; r < FIR_R; r++) w[r] = Us[r];
        int sidx = 0, k = k0;
        for (; k + (FIR_R - 1) * Q < D; k += FIR_R * Q) {
#pragma unroll
            for (int t = 0; t < FIR_R; t++) {
                const double cf = __ldg(T + D + k + t * Q);
#pragma unroll
                for (int r = 0; r < FIR_R; r++) acc[r] = fma(cf, w[(r + t) % FIR_R], acc[r]);
                sidx++;
                w[t] = Us[FIR_R - 1 + sidx];
            }
        }
        for (; k < D; k += Q) {
            const double cf = __ldg(T + D + k);
#pragma unroll
            for (int r = 0; r < FIR_R; r++) acc[r] = fma(cf, w[r], acc[r]);
#pragma unroll
            for (int r = 0; r < FIR_R - 1; r++) w[r] = w[r + 1];
            sidx++;
            w[FIR_R - 1] = Us[FIR_R - 1 + sidx];
        }
    }
#pragma unroll
    for (int r = 0; r < FIR_R; r++) {
        long long m = m0 + (long long)lane * FIR_R + r;
        long long j = (long long)P * m + ph + 1;
        if (j <= J.nout) out[J.out_off + j - 1] = acc[r];
    }
}

// Outputs the regular pattern does not describe (depth clipped at the ends of the sound, an index that is an exact
// integer, a fractional position that fell on the other side of an integer than the table's) are recomputed exactly.
__global__ void __launch_bounds__(256) k_sinc_fixup(const ResampleJob* __restrict__ jobs, const long long* __restrict__ out_prefix,
                                                     int njobs, const double* __restrict__ filt, const double* __restrict__ table_fl,
                                                     const long long* __restrict__ table_mid, double* __restrict__ out, int P,
                                                     int Q, int D, double dx_src) {
    const int lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total = out_prefix[njobs];
    for (long long o0 = gw * 32; o0 < total; o0 += nw * 32) {
        const long long o = o0 + lane;
        bool irregular = false;
        int job = 0;
        long long j = 0;
        double index = 0.0;
        if (o < total) {
            job = find_segment_ll(out_prefix, njobs, o);
            const ResampleJob J = jobs[job];
            j = o - out_prefix[job] + 1;
            const double x = J.out_x1 + (double)(j - 1) * J.out_dx;
            index = (x - J.x1) / dx_src + 1.0;
            const long long midleft = (long long)floor(index);
            const int ph = (int)((j - 1) % P);
            const long long m = (j - 1) / P, m_rep = J.nout / 2 / P;
            const long long pred = table_mid[(size_t)J.table_id * P + ph] + (long long)Q * (m - m_rep);
            const double tfl = table_fl[(size_t)J.table_id * P + ph];
            irregular = midleft != pred || midleft < D || midleft + D > J.nx || index == (double)midleft ||
                        !(fabs((index - (double)midleft) - tfl) < 1e-8);
        }
        unsigned mask = __ballot_sync(FULL_MASK, irregular);
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const int jb = __shfl_sync(FULL_MASK, job, src);
            const long long jj = __shfl_sync(FULL_MASK, j, src);
            const double idx = __shfl_sync(FULL_MASK, index, src);
            const ResampleJob J = jobs[jb];
            double v = sinc_interp_warp_ll(filt + J.filt_off - 1, J.nx, idx, D, lane);
            if (lane == 0) out[J.out_off + jj - 1] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ host driver
static void plan_passes(int logn, int* cb, int* npass, int rbits[4]) {
    *cb = logn < 12 ? logn : 12;
    int rem = logn - *cb;
    *npass = 0;
    if (rem <= 0) return;
    int np = (rem + 7) / 8;
    for (int i = 0; i < np; i++) { rbits[i] = rem / np + (i < rem % np ? 1 : 0); }
    *npass = np;
}

// jobs of one FFT size: d_ids lists the job indices, cnt of them
void launch_resample_fft_group(const ResampleJob* d_jobs, const int* d_ids, int cnt, int logn, SPtr pcm, double2* zbuf,
                               double* filt, const double2* tw, double upfactor, cudaStream_t s, long long* launches, int mode) {
    int cb, npass, rbits[4];
    plan_passes(logn, &cb, &npass, rbits);
    const long long N = 1LL << logn;
    if (npass == 0) {
        size_t smem = sizeof(double2) * (1u << cb);
        cudaFuncSetAttribute(k_fft_inner<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_fft_inner<true><<<cnt, RS_NTHR, smem, s>>>(d_jobs, d_ids, pcm, zbuf, filt, tw, logn, cb, upfactor, mode);
        (*launches)++;
        return;
    }
    // forward strided passes (outermost first)
    int nl = logn;
    for (int i = 0; i < npass; i++) {
        int rb = rbits[i];
        size_t smem = sizeof(double2) * ((size_t)1 << rb) * TB;
        long long tiles = N / (((long long)1 << rb) * TB);
        unsigned grid = (unsigned)(tiles * cnt);
        if (i == 0) {
            cudaFuncSetAttribute(k_fft_strided<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_fft_strided<false, true, false><<<grid, RS_NTHR, smem, s>>>(d_jobs, d_ids, pcm, zbuf, filt, tw, logn, nl, rb, mode);
        } else {
            cudaFuncSetAttribute(k_fft_strided<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_fft_strided<false, false, false><<<grid, RS_NTHR, smem, s>>>(d_jobs, d_ids, pcm, zbuf, filt, tw, logn, nl, rb, mode);
        }
        (*launches)++;
        nl -= rb;
    }
    {
        size_t smem = sizeof(double2) * (1u << cb);
        unsigned grid = (unsigned)((N >> cb) * cnt);
        cudaFuncSetAttribute(k_fft_inner<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_fft_inner<false><<<grid, RS_NTHR, smem, s>>>(d_jobs, d_ids, pcm, zbuf, filt, tw, logn, cb, upfactor, mode);
        (*launches)++;
    }
    // inverse strided passes (innermost first)
    for (int i = npass - 1; i >= 0; i--) {
        int rb = rbits[i];
        nl += rb;
        size_t smem = sizeof(double2) * ((size_t)1 << rb) * TB;
        long long tiles = N / (((long long)1 << rb) * TB);
        unsigned grid = (unsigned)(tiles * cnt);
        if (i == 0) {
            cudaFuncSetAttribute(k_fft_strided<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_fft_strided<true, false, true><<<grid, RS_NTHR, smem, s>>>(d_jobs, d_ids, pcm, zbuf, filt, tw, logn, nl, rb, mode);
        } else {
            cudaFuncSetAttribute(k_fft_strided<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_fft_strided<true, false, false><<<grid, RS_NTHR, smem, s>>>(d_jobs, d_ids, pcm, zbuf, filt, tw, logn, nl, rb, mode);
        }
        (*launches)++;
    }
}

void launch_sinc_resample(const ResampleJob* d_jobs, const long long* d_out_prefix, int njobs, long long total_out_hint,
                          const int* d_table_rep, int ntables, const int* d_tile_prefix, int total_tiles, const double* filt,
                          double* table, double* out, int P, int Q, int D, double dx_src, cudaStream_t s, long long* launches) {
    // P <= 16 phases with Q <= 64: polyphase FIR kernel; more phases (44.1 kHz -> 16 kHz has 160): coefficient rows per phase
    // but one warp per output sample; P == 0 (no short rational period): every sample evaluates the window itself.
    const bool fir = P > 0 && P <= 16 && Q <= 64;
    double* table_fl = table + (size_t)ntables * (P > 0 ? P : 1) * 2 * D;      // stored behind the coefficient rows
    long long* table_mid = (long long*)(table_fl + (size_t)ntables * (P > 0 ? P : 1));
    long long blocks = (total_out_hint + 255) / 256;
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    if (blocks < 1) blocks = 1;
    if (fir && ntables > 0 && total_tiles > 0) {
        k_sinc_table<<<ntables, 256, 0, s>>>(d_jobs, d_table_rep, ntables, table, table_fl, table_mid, P, D, dx_src);
        const int span = Q * (32 * FIR_R - 1) + 2 * D + 3 * Q + Q;               // upper bound of the tile span
        const size_t smem = sizeof(double) * (size_t)Q * ((span + Q - 1) / Q + 4);
        cudaFuncSetAttribute(k_sinc_fir, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_sinc_fir<<<total_tiles, P * 32, smem, s>>>(d_jobs, d_tile_prefix, njobs, filt, table, table_mid, out, P, Q, D);
        k_sinc_fixup<<<(unsigned)blocks, 256, 0, s>>>(d_jobs, d_out_prefix, njobs, filt, table_fl, table_mid, out, P, Q, D, dx_src);
        (*launches) += 3;
    } else {
        if (P > 0 && ntables > 0) {
            k_sinc_table<<<ntables, 256, 0, s>>>(d_jobs, d_table_rep, ntables, table, table_fl, table_mid, P, D, dx_src);
            (*launches)++;
        }
        k_sinc_apply<<<(unsigned)blocks, 256, 0, s>>>(d_jobs, d_out_prefix, njobs, filt, table, table_fl, out, ntables > 0 ? P : 0, D, dx_src);
        (*launches)++;
    }
}
