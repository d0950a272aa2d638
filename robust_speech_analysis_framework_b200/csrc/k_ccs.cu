// k_ccs.cu -- forward cross-correlation pitch frames (Sound_into_PitchFrame, FCC_NORMAL branch of fon/Sound_to_Pitch.cpp),
// the arithmetic behind to_harmonicity_cc at mshds_extractor.py:221 and to_pitch_cc at :320.
//
// Praat correlates every frame on its own: sum_j x[j] x[j + lag] over a window of W samples for every lag up to
// maximumLag (W = 1198, 268 lags for the harmonicity analysis of a low-pitched speaker: 321,000 multiply-adds per frame),
// and the frames are only 80 samples (5 ms) apart, so consecutive windows share 93 % of their products.  This kernel
// computes the products ONCE, in blocks of one frame step:
//
//   P_b[lag] = sum over the H samples j of block b of s[j] s[j + lag]            (H = dt / dx = 80; lattice anchored at the
//                                                                                 clip's first frame)
//   C[lag]   = sum of the P_b of the blocks inside the frame's window  +/-  the few products at the two window edges that the
//              lattice does not describe (2 samples per frame for the configurations of this path)
//
// and obtains the mean-subtracted correlation of the reference from raw-sample sums,
//   sum (s[j] - m)(s[j + lag] - m) = C[lag] - m (A_0 + A_lag) + W m^2,     A_lag = sum_j s[j + lag],
// with A_lag and the normalisers sum (s - m)^2 from running prefix sums of s and s^2.  The recordings are int16: every
// product is an integer multiple of 2^-30 below 2^53, so P_b, C, A and the prefix sums are EXACT in float64 whatever the
// summation order -- the result does not depend on how the frames are grouped into runs, and it is closer to the true
// correlation than the reference's own 1,198 rounded additions.  (float64 input behind the resampling front-end keeps the
// frame-by-frame kernel k_pitch_frames<true>: there the grouping would show in the last bits.)
//
// Work per frame drops from W x maximumLag to (H + edge) x maximumLag multiply-adds + one pass over the ring: 13x fewer for
// the harmonicity pass, 3x for the one-period window of to_pitch_cc.
//
// Two kernels share this derivation.  The default is k_cc_frames_w further down (one WARP per run of frames, the block sums
// replaced by sliding sums in registers, no block barrier: 2.2x faster than the kernel described next); k_cc_frames_s, the first
// implementation, is kept behind the development switch "legacy_cc" = 2 as an independent checker -- both are exact, so their
// rows are bit-identical (tests/test_gpu_parity.py::test_shared_block_cross_correlation_equals_frame_by_frame).
//
// Structure of k_cc_frames_s: a CTA walks a run of consecutive frames of one recording.  Per step (= one new block): warps 0-3 form the
// block's products (a thread owns 13 consecutive lags and a slice of the block; the lag window slides through registers as
// a ring: 2 shared-memory loads per 13 DFMA), warp 4 meanwhile loads the next H samples and extends the prefix sums; then
// all threads add the partial products into a ring of the last R blocks and assemble the frame that became complete
// (sum over the ring, edges, normalisation), and the frame's local peak, maxima (harmonicity: queued for k_hnr_refine) or
// hand-over flags (to_pitch_cc: candidates by k_ac_candidates) follow while the next block is already being multiplied.
#include <cstdlib>
#include "internal.h"
#include "common.cuh"

#define CCS_PW 4                        // product warps
#define CCS_NTP (32 * CCS_PW)
#define CCS_NT (CCS_NTP + 32)           // + the loader warp
#define CCS_TL 13                       // lags per thread (odd: conflict-free lag windows)
#define CCS_MIRROR 512                  // the first CCS_MIRROR samples of the circular buffer are repeated behind its end

struct CcsParams {
    const int16_t* pcm;
    int H;                  // frame step in samples
    int R;                  // ring slots (blocks kept)
    int cap;                // circular capacity of the sample / prefix buffers (power of two)
    int LS;                 // doubles per ring / partial row
    int nchunk_max;
    int run;                // frames per turn
    int o_x, o_ps, o_pq, o_ring, o_part, o_rr;      // byte offsets into dynamic shared memory
};

// frame k of a clip: first sample of the correlation window, usable lags, left sample of the frame centre
__device__ __forceinline__ void ccs_geom(const PitchCfg& g, double x1, double dx, double t1, int k, long long nx,
                                         long long& lo, int& Lloc, long long& left) {
    const double t = t1 + (double)k * g.dt;
    left = x_to_low(x1, dx, t);
    const double startTime = t - 0.5 * (1.0 / g.floor_hz + g.dt_window);
    long long startS = x_to_low(x1, dx, startTime);
    if (startS < 1) startS = 1;
    long long localSpan = g.maximumLag + g.nsamp_window;
    if (localSpan > nx + 1 - startS) localSpan = nx + 1 - startS;
    const int l = (int)(localSpan - g.nsamp_window);
    Lloc = l > 0 ? l : 0;
    lo = startS;
}

// floor(a / b) for 32-bit offsets from the lattice origin (64-bit integer division is a ~150-instruction routine)
__device__ __forceinline__ int ccs_floordiv(long long a, int b) { const int x = (int)a; return x >= 0 ? x / b : -((-x + b - 1) / b); }

// loader warp: samples [from, to) of the clip (1-based; zero outside the clip) into the circular buffer, prefix sums extended
__device__ __forceinline__ void ccs_extend(const int16_t* __restrict__ pcm, long long nx, long long from, long long to, int lane,
                                           double* xb, double* ps, double* pq, int mask, int cap, double& totS, double& totQ) {
    const int n = (int)(to - from);
    if (n <= 0) return;
    const int per = (n + 31) >> 5;
    const long long i0 = from + (long long)lane * per;
    long long i1 = i0 + per;
    if (i1 > to) i1 = to;
    double ls = 0.0, lq = 0.0;
    if (per <= 4) {                               // one frame step: all loads of the lane are in flight together
        short raw[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u;
            raw[u] = (i < i1 && i >= 1 && i <= nx) ? __ldg(pcm + i - 1) : (short)0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u;
            if (i < i1) {
                const double v = (double)raw[u] * (1.0 / 32768.0);
                const int pos = (int)(i & (long long)mask);
                xb[pos] = v;
                if (pos < CCS_MIRROR) xb[cap + pos] = v;
                ls += v;
                lq = fma(v, v, lq);
            }
        }
    } else {
        for (long long i = i0; i < i1; i++) {
            const double v = (i >= 1 && i <= nx) ? (double)__ldg(pcm + i - 1) * (1.0 / 32768.0) : 0.0;
            const int pos = (int)(i & (long long)mask);
            xb[pos] = v;
            if (pos < CCS_MIRROR) xb[cap + pos] = v;
            ls += v;
            lq = fma(v, v, lq);
        }
    }
    double is = ls, iq = lq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double a = __shfl_up_sync(FULL_MASK, is, o), b = __shfl_up_sync(FULL_MASK, iq, o);
        if (lane >= o) { is += a; iq += b; }
    }
    double rs = totS + (is - ls), rq = totQ + (iq - lq);
    for (long long i = i0; i < i1; i++) {
        const int pos = (int)(i & (long long)mask);
        const double v = xb[pos];
        rs += v;
        rq = fma(v, v, rq);
        ps[pos] = rs;
        pq[pos] = rq;
    }
    totS += __shfl_sync(FULL_MASK, is, 31);
    totQ += __shfl_sync(FULL_MASK, iq, 31);
}

// What follows the correlation of one frame (executed by ONE warp; rr[0 .. maximumLag + 1] = the frame's r in shared memory):
// relative local peak, then either the frame's maxima as items for k_hnr_refine (harmonicity) or the hand-over flag for
// k_ac_candidates (to_pitch_cc).
__device__ __forceinline__ void ccs_frame_tail(const Clips& c, const PitchPass& p, const PitchCfg& g, int clip, int fidx,
                                               const double* rr, double localPeak, double gpeak, int Lm, double dx, int lane) {
    const double intensity = localPeak > gpeak ? 1.0 : localPeak / gpeak;
    if (p.hnr_mode) {
        double uvs = g.sil <= 0 ? 0.0 : 2.0 - intensity / (g.sil / (1.0 + g.vt));
        uvs = g.vt + (uvs > 0 ? uvs : 0);
        if (localPeak != 0.0 && uvs < 1.0) {
            // every maximum of r is an item for k_hnr_refine: count them, reserve the queue slots with ONE
            // atomic per frame (a single address takes the atomics of the whole grid), then write the items
            const double thr = 0.5 * g.vt;
            const int B = g.brent_ixmax;
            const int upper = Lm < B ? Lm : B;
            int count = 0;
            for (int i0 = 2; i0 < upper; i0 += 32) {
                const int i = i0 + lane;
                bool flag = false;
                if (i < upper) { const double ri = rr[i]; flag = ri > thr && ri > rr[i - 1] && ri >= rr[i + 1]; }
                count += __popc(__ballot_sync(FULL_MASK, flag));
            }
            if (count > 0) {
                unsigned long long q0 = 0;
                if (lane == 0) q0 = atomicAdd(p.qcount64, (unsigned long long)count);
                q0 = __shfl_sync(FULL_MASK, q0, 0);
                for (int i0 = 2; i0 < upper; i0 += 32) {
                    const int i = i0 + lane;
                    bool flag = false;
                    double ri = 0.0, rm = 0.0, rp = 0.0;
                    if (i < upper) { ri = rr[i]; rm = rr[i - 1]; rp = rr[i + 1]; flag = ri > thr && ri > rm && ri >= rp; }
                    const unsigned m = __ballot_sync(FULL_MASK, flag);
                    if (flag) {
                        const double dr = 0.5 * (rp - rm), d2r = 2 * ri - rm - rp;
                        const double freq = 1.0 / dx / (i + dr / d2r);
                        const unsigned long long item = ((unsigned long long)(unsigned)fidx << 32) |
                                                        ((unsigned long long)i << 8) | (freq > 0.3 / dx ? 1ull : 0ull);
                        const unsigned long long q = q0 + (unsigned long long)__popc(m & ((1u << lane) - 1u));
                        if (q < p.q64_cap) p.queue64[q] = item;
                        else atomicOr(&c.status[clip], ST_HNR);      // cannot happen: capacity is the worst case
                    }
                    q0 += (unsigned long long)__popc(m);
                }
            }
        }
        if (lane == 0) { p.inten[fidx] = intensity; p.best_bits[fidx] = 0ull; }
    } else if (lane == 0) {
        p.inten[fidx] = intensity;
        p.ncand[fidx] = localPeak != 0.0 ? 1 : 0;        // hand-over to k_ac_candidates
    }
}

// products of one block: part[ch][lag - 1] = sum over slice ch of the block of x[j] x[j + lag], lag = 1..ngroups*TL
template <int TL>
__device__ __forceinline__ void ccs_products(const double* __restrict__ xa0, double* __restrict__ part, int LS, int H, int ngroups,
                                             int nchunk, int tid) {
    const int q = (H + nchunk - 1) / nchunk;
    for (int wi = tid; wi < ngroups * nchunk; wi += CCS_NTP) {
        const int grp = wi % ngroups, ch = wi / ngroups;
        const int lag0 = 1 + TL * grp;
        const int j0 = ch * q;
        int len = (j0 + q < H ? j0 + q : H) - j0;
        if (len < 0) len = 0;
        double acc[TL], yw[TL];
        const double* xa = xa0 + j0;
        const double* xl = xa0 + j0 + lag0;
#pragma unroll
        for (int u = 0; u < TL; u++) { acc[u] = 0.0; yw[u] = xl[u]; }
        int j = 0;
        // TL steps per trip, the lag window kept as a register ring (sample xl[k] sits in slot k % TL)
        for (; j + TL <= len; j += TL) {
#pragma unroll
            for (int t = 0; t < TL; t++) {
                const double xv = xa[j + t];
#pragma unroll
                for (int u = 0; u < TL; u++) acc[u] = fma(xv, yw[(u + t) % TL], acc[u]);
                yw[t] = xl[j + t + TL];
            }
        }
        if (j < len) {
#pragma unroll
            for (int t = 0; t < TL; t++) {
                if (j + t < len) {
                    const double xv = xa[j + t];
#pragma unroll
                    for (int u = 0; u < TL; u++) acc[u] = fma(xv, yw[(u + t) % TL], acc[u]);
                    yw[t] = xl[j + t + TL];
                }
            }
        }
        double* pr = part + (size_t)ch * LS + (lag0 - 1);
#pragma unroll
        for (int u = 0; u < TL; u++) pr[u] = acc[u];
    }
}

__global__ void __launch_bounds__(CCS_NT, 2) k_cc_frames_s(const __grid_constant__ Clips c, const __grid_constant__ PitchPass p,
                                                           const __grid_constant__ CcsParams A) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* xb = (double*)(smem + A.o_x);
    double* ps = (double*)(smem + A.o_ps);
    double* pq = (double*)(smem + A.o_pq);
    double* ring = (double*)(smem + A.o_ring);
    double* part = (double*)(smem + A.o_part);
    double* rr = (double*)(smem + A.o_rr);
    __shared__ double s_red[CCS_PW + 1];
    __shared__ int s_turn, s_clip;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = p.fstart[c.n];
    const int nturn = (total + A.run - 1) / A.run;
    const int mask = A.cap - 1, cap = A.cap, H = A.H, R = A.R, LS = A.LS;
    const double dx = c.dx;
    double totS = 0.0, totQ = 0.0;              // loader warp: running prefix totals of the current run

    for (;;) {
        __syncthreads();
        if (tid == 0) s_turn = atomicAdd(p.turn_counter, 1);
        __syncthreads();
        const int turn = s_turn;
        if (turn >= nturn) break;
        int f = turn * A.run;
        const int f_end = f + A.run < total ? f + A.run : total;
        while (f < f_end) {
            __syncthreads();
            if (tid == 0) s_clip = find_segment(p.fstart, c.n, f);
            __syncthreads();
            const int clip = s_clip;
            const int cend = p.fstart[clip + 1];
            const int nseg = (f_end < cend ? f_end : cend) - f;
            const int k0 = f - p.fstart[clip];
            const PitchCfg& g = p.cfg[c.cls[clip]];
            const int W = g.nsamp_window, Lm = g.maximumLag, np = g.nsamp_period, hp = g.halfnsamp_period, hw = g.halfnsamp_window;
            const int Ls = stored_lags(g);
            const long long base = c.off[clip], nx = c.off[clip + 1] - base;
            const int16_t* pcm = A.pcm + base;
            const double x1 = c.x1[clip], t1 = p.t1[clip];
            const double gpeak = c.gpeak[clip];
            const int ngroups = (Lm + CCS_TL - 1) / CCS_TL;
            int nchunk = CCS_NTP / ngroups;
            if (nchunk < 1) nchunk = 1;
            if (nchunk > A.nchunk_max) nchunk = A.nchunk_max;
            double* rrow0 = p.rbuf + (size_t)f * p.rstride;

            long long G0, lo, left;
            int Lloc;
            ccs_geom(g, x1, dx, t1, 0, nx, G0, Lloc, left);                 // lattice origin: first frame of the CLIP
            ccs_geom(g, x1, dx, t1, k0, nx, lo, Lloc, left);
            int b = ccs_floordiv(lo - G0 + H / 2, H);                       // first block of this run
            {
                long long o = lo;
                if (left + 1 - np < o) o = left + 1 - np;
                if (G0 + (long long)H * b < o) o = G0 + (long long)H * b;
                if (warp == CCS_PW) {
                    totS = 0.0; totQ = 0.0;
                    if (lane == 0) { ps[(int)((o - 1) & (long long)mask)] = 0.0; pq[(int)((o - 1) & (long long)mask)] = 0.0; }
                    __syncwarp();
                    ccs_extend(pcm, nx, o, G0 + (long long)H * (b + 1) + Lm, lane, xb, ps, pq, mask, cap, totS, totQ);
                }
            }
            int slot = (int)(b % R);
            int kk = 0;
            __syncthreads();
            while (kk < nseg) {
                // ---- phase B: products of block b | next samples
                if (warp < CCS_PW) {
                    const long long gs = G0 + (long long)H * b;
                    ccs_products<CCS_TL>(xb + (int)(gs & (long long)mask), part, LS, H, ngroups, nchunk, tid);
                } else {
                    const long long fr = G0 + (long long)H * (b + 1) + Lm;
                    ccs_extend(pcm, nx, fr, fr + H, lane, xb, ps, pq, mask, cap, totS, totQ);
                }
                __syncthreads();
                // ---- phase S: the block enters the ring
                // (thread t owns lags t, t + CCS_NT, ... here AND in the assembly below: no barrier in between)
                for (int l = tid; l <= Lm; l += CCS_NT) {
                    if (l == 0) continue;
                    double v = 0.0;
                    for (int ch = 0; ch < nchunk; ch++) v += part[(size_t)ch * LS + l - 1];
                    ring[(size_t)slot * LS + l - 1] = v;
                }
                // ---- frames that are complete with block b (normally one)
                bool first = true;
                for (;;) {
                    if (kk >= nseg) break;
                    const long long hi = lo + W;
                    long long need = hi;                                              // samples needed, less the lag reach
                    if (left + np + 1 - Lm > need) need = left + np + 1 - Lm;
                    if (left + hp + 1 - Lm > need) need = left + hp + 1 - Lm;
                    const int e = ccs_floordiv(need - G0 + H - 1, H) - 1;             // step with which the frame is complete
                    if (e > b) break;
                    if (!first) __syncthreads();                                      // rr / s_red of the previous frame are consumed
                    first = false;
                    const int bA = ccs_floordiv(lo - G0 + H / 2, H), bB = ccs_floordiv(hi - G0 + H / 2, H);
                    const long long latA = G0 + (long long)H * bA, latB = G0 + (long long)H * bB;
                    const int fidx = f + kk;
                    double* rrow = rrow0 + (size_t)kk * p.rstride;
                    // local mean (exact sums) and the raw-sample sums of the window
#define PSV(i) ps[(int)((i) & (long long)mask)]
#define PQV(i) pq[(int)((i) & (long long)mask)]
#define XV(i) xb[(int)((i) & (long long)mask)]
                    const double mu = (PSV(left + np) - PSV(left - np)) / (double)(2 * np);
                    const double A0 = PSV(hi - 1) - PSV(lo - 1), Q0 = PQV(hi - 1) - PQV(lo - 1);
                    const double wmm = (double)W * mu * mu;
                    const double sx = fma(-2.0 * mu, A0, Q0) + wmm;
                    const int sA = (int)(bA % R);
                    for (int l = tid; l < Ls; l += CCS_NT) {
                        double v = 0.0;
                        if (l == 0) v = 1.0;
                        else if (l <= Lloc) {
                            double C = 0.0;
                            int s2 = sA;
                            for (int bb = bA; bb < bB; bb++) { C += ring[(size_t)s2 * LS + l - 1]; s2 = s2 + 1 == R ? 0 : s2 + 1; }
                            if (lo > latA) { for (long long j = latA; j < lo; j++) C = fma(-XV(j), XV(j + l), C); }
                            else { for (long long j = lo; j < latA; j++) C = fma(XV(j), XV(j + l), C); }
                            if (hi > latB) { for (long long j = latB; j < hi; j++) C = fma(XV(j), XV(j + l), C); }
                            else { for (long long j = hi; j < latB; j++) C = fma(-XV(j), XV(j + l), C); }
                            const double Al = PSV(hi - 1 + l) - PSV(lo - 1 + l), Ql = PQV(hi - 1 + l) - PQV(lo - 1 + l);
                            const double pr = fma(-mu, A0 + Al, C) + wmm;
                            const double sy = fma(-2.0 * mu, Al, Ql) + wmm;
                            v = pr / sqrt(sx * sy);
                        }
                        if (l <= Lm + 1) rr[l] = v;
                        rrow[l] = v;
                    }
                    if (tid == 0 && Ls <= Lm + 1) rr[Lm + 1] = 0.0;
                    // local peak: max |s - mean| over the middle of the (centred) analysis window
                    {
                        const long long right = left + 1;
                        long long a0 = right - hp, a1 = right + hp - 1;
                        if (a0 < right - hw) a0 = right - hw;
                        if (a1 > right + hw - 1) a1 = right + hw - 1;
                        double lp = 0.0;
                        for (long long i = a0 + tid; i <= a1; i += CCS_NT) lp = fmax(lp, fabs(XV(i) - mu));
                        lp = warp_max(lp);
                        if (lane == 0) s_red[warp] = lp;
                    }
                    __syncthreads();
                    // ---- phase D (loader warp only; the product warps go on to the next block): local peak, maxima / hand-over
                    if (warp == CCS_PW) {
                        double localPeak = s_red[0];
#pragma unroll
                        for (int w = 1; w <= CCS_PW; w++) localPeak = fmax(localPeak, s_red[w]);
                        ccs_frame_tail(c, p, g, clip, fidx, rr, localPeak, gpeak, Lm, dx, lane);
                    }
#undef PSV
#undef PQV
#undef XV
                    kk++;
                    if (kk < nseg) ccs_geom(g, x1, dx, t1, k0 + kk, nx, lo, Lloc, left);
                }
                if (first) __syncthreads();          // no frame this step: part[] is reused by the next block's products
                b++;
                slot = slot + 1 == R ? 0 : slot + 1;
            }
            f += nseg;
        }
    }
}

// ================================================================================================ warp-per-run version
// k_cc_frames_w: the same exact arithmetic with the block ring replaced by SLIDING sums held in registers, and the CTA-wide
// pipeline (five warps, three block barriers per frame: ncu showed 37 % of the stall samples of k_cc_frames_s on the barrier
// and the FP64 pipe at 10 %) replaced by ONE WARP PER RUN of consecutive frames with no block barrier at all:
//
//   C_k+1[lag] = C_k[lag] + sum_{j in [hi_k, hi_k+1)} s[j] s[j + lag] - sum_{j in [lo_k, lo_k+1)} s[j] s[j + lag]
//
// (window [lo, hi) of frame k, hi = lo + W).  Every term is an integer multiple of 2^-30 far below 2^53, so adding and
// subtracting is exact: no drift, the result does not depend on where a run starts, and it is bit-identical to
// k_cc_frames_s.  A lane owns TL consecutive lags (TL = 5, 7, 9: 32 TL >= maximumLag + 1) with their running sums in
// registers; per frame it forms 2 x (frame step) x TL products from two short sample windows in shared memory (the lag
// window slides through a register ring: one broadcast load + one conflict-free load per TL DFMA).  The raw-sample sums of
// the normalisation follow the same way: A_0 and sum s^2 slide with integer warp reductions, A_lag = A_0 + a prefix scan of
// s[hi + m - 1] - s[lo + m - 1] over the lags (lane-local sums + one warp scan); the local mean slides as an integer sum,
// the local peak is an integer min / max.  The first frame of a run is built by sliding an empty window open (W / 96 block
// additions ~ the cost of 7 ordinary frames).
#define CCW_WARPS 8
#define CCW_NT (32 * CCW_WARPS)
#define CCW_DMAX 96                              // largest window advance (and start-up block) in samples
#define CCW_XW (CCW_DMAX + 32 * 9 + 8)           // doubles per edge window
#define CCW_RR 352                               // doubles of the r row copy (lags 0 .. 32 TL)
#define CCW_WARP_DOUBLES (2 * CCW_XW + CCW_RR)

struct CcwParams {
    const int16_t* pcm;
    int run;                // frames per turn
};

// The straight-line code a warp executes per frame has to stay well inside the 32 KB instruction cache (the first version
// inlined the loaders, the products and nine copies of the division / square root at every call site and for every TL: 13,600
// instructions, 35 % of its stall samples waiting for instruction fetch).  Hence: helpers that do not touch the register
// arrays are __noinline__, the products have ONE call site inside a two-trip loop, and the normalisation runs as a rolled
// loop over rows passed through shared memory.

// samples [base, base + count) of the clip (1-based, zero outside) -> xw[0 .. count) as float64; integer sum and sum of squares of
// the first blk (<= 96) of them
__device__ __noinline__ void ccw_load(const int16_t* __restrict__ pcm, int nx, int base, int count, int blk, int xw_off /* doubles */,
                                      int& s_out, long long& q_out) {
    extern __shared__ __align__(16) unsigned char ccw_smem[];
    double* xw = (double*)ccw_smem + xw_off;
    const int lane = threadIdx.x & 31;
    int ps = 0;
    unsigned pq = 0;
    for (int e0 = 0; e0 < count; e0 += 128) {
        int v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + 32 * u + lane, i = base + e;
            v[u] = (e < count && i >= 1 && i <= nx) ? (int)__ldg(pcm + i - 1) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int e = e0 + 32 * u + lane;
            if (e < count) xw[e] = (double)v[u] * (1.0 / 32768.0);
            if (e < blk) { ps += v[u]; pq += (unsigned)(v[u] * v[u]); }
        }
    }
    s_out = __reduce_add_sync(FULL_MASK, ps);
    const unsigned ql = __reduce_add_sync(FULL_MASK, pq & 0xffffu), qh = __reduce_add_sync(FULL_MASK, pq >> 16);
    q_out = ((long long)qh << 16) + (long long)ql;
}

// integer sum of the clip samples a .. b (1-based, inclusive, zero outside the clip), minus those of a2 .. b2: one reduction
__device__ __noinline__ int ccw_isum2(const int16_t* __restrict__ pcm, int nx, int a, int b, int a2, int b2) {
    const int lane = threadIdx.x & 31;
    int acc = 0;
    for (int i = a + lane; i <= b; i += 32) acc += (i >= 1 && i <= nx) ? (int)__ldg(pcm + i - 1) : 0;
    for (int i = a2 + lane; i <= b2; i += 32) acc -= (i >= 1 && i <= nx) ? (int)__ldg(pcm + i - 1) : 0;
    return __reduce_add_sync(FULL_MASK, acc);
}

// max |s - mu| over the clip samples a0 .. a1 from their integer extremes (rounding is monotone: the extremes give the maximum)
__device__ __noinline__ double ccw_peak(const int16_t* __restrict__ pcm, int nx, int a0, int a1, double mu) {
    const int lane = threadIdx.x & 31;
    int vmin = 32767, vmax = -32768;
    for (int i = a0 + lane; i <= a1; i += 32) {
        const int v = (i >= 1 && i <= nx) ? (int)__ldg(pcm + i - 1) : 0;
        vmin = v < vmin ? v : vmin;
        vmax = v > vmax ? v : vmax;
    }
    vmin = __reduce_min_sync(FULL_MASK, vmin);
    vmax = __reduce_max_sync(FULL_MASK, vmax);
    return a1 >= a0 ? fmax(fabs((double)vmax * (1.0 / 32768.0) - mu), fabs((double)vmin * (1.0 / 32768.0) - mu)) : 0.0;
}

// r[l] of lags 1 .. NL from the rows C (in xn), A_lag (in xo), Q_lag (in rr) -> rr and the frame's row in HBM; then the tail
__device__ __noinline__ void ccw_finish(const Clips& c, const PitchPass& p, const PitchCfg& g, const int16_t* __restrict__ pcm, int nx,
                                        int clip, int fidx, int left, int Lloc, int NL, double mu, double A0, double Q0, int warp_off) {
    extern __shared__ __align__(16) unsigned char ccw_smem[];
    double* xn = (double*)ccw_smem + warp_off;
    double* xo = xn + CCW_XW;
    double* rr = xo + CCW_XW;
    const int lane = threadIdx.x & 31;
    const int W = g.nsamp_window, Lm = g.maximumLag, hp = g.halfnsamp_period, hw = g.halfnsamp_window;
    const int Ls = stored_lags(g);
    const double wmm = (double)W * mu * mu;
    const double sx = fma(-2.0 * mu, A0, Q0) + wmm;
    double* rrow = p.rbuf + (size_t)fidx * p.rstride;
#pragma unroll 1
    for (int l = 1 + lane; l <= NL; l += 32) {
        double v = 0.0;
        if (l <= Lloc) {
            const double Al = xo[l], Ql = rr[l];
            const double pr = fma(-mu, A0 + Al, xn[l]) + wmm;
            const double sy = fma(-2.0 * mu, Al, Ql) + wmm;
            v = pr / sqrt(sx * sy);
        }
        rr[l] = v;
        if (l < Ls) rrow[l] = v;
    }
    if (lane == 0) { rr[0] = 1.0; rrow[0] = 1.0; }
    for (int l = NL + 1 + lane; l < Ls; l += 32) rrow[l] = 0.0;
    // local peak: max |s - mean| over the middle of the (centred) analysis window
    const int right = left + 1;
    int a0 = right - hp, a1 = right + hp - 1;
    if (a0 < right - hw) a0 = right - hw;
    if (a1 > right + hw - 1) a1 = right + hw - 1;
    const double localPeak = ccw_peak(pcm, nx, a0, a1, mu);
    __syncwarp();
    ccs_frame_tail(c, p, g, clip, fidx, rr, localPeak, c.gpeak[clip], Lm, c.dx, lane);
}

// C[u] += sum_{t < len} x[t] x[t + lag], lag = 1 + TL lane + u; xw[e] = sample (block start + e)
template <int TL>
__device__ __forceinline__ void ccw_products(const double* __restrict__ xw, int len, int lane, double (&C)[TL]) {
    const double* xl = xw + 1 + TL * lane;
    double yw[TL];
#pragma unroll
    for (int u = 0; u < TL; u++) yw[u] = xl[u];
    int t0 = 0;
#pragma unroll 1
    for (; t0 + TL <= len; t0 += TL) {
#pragma unroll
        for (int t = 0; t < TL; t++) {
            const double xv = xw[t0 + t];
#pragma unroll
            for (int u = 0; u < TL; u++) C[u] = fma(xv, yw[(u + t) % TL], C[u]);
            yw[t] = xl[t0 + t + TL];
        }
    }
    // the last len % TL samples: the ring is dead afterwards, so it can be read at a run-time rotation through shared memory ...
    // simpler: one more (partly predicated) trip
#pragma unroll
    for (int t = 0; t < TL - 1; t++) {
        if (t0 + t < len) {
            const double xv = xw[t0 + t];
#pragma unroll
            for (int u = 0; u < TL; u++) C[u] = fma(xv, yw[(u + t) % TL], C[u]);
            yw[t] = xl[t0 + t + TL];
        }
    }
}

// nseg consecutive frames (k0 ..) of one clip, flat frame index f ..: executed by one warp
template <int TL>
__device__ __noinline__ void ccw_segment(const Clips& c, const PitchPass& p, const int16_t* __restrict__ pcm_all, int clip, int f,
                                         int nseg, int k0) {
    extern __shared__ __align__(16) unsigned char ccw_smem[];     // the kernel's dynamic shared memory: one region per warp
    const int lane = threadIdx.x & 31;
    const int warp_off = (threadIdx.x >> 5) * CCW_WARP_DOUBLES;
    double* xn = (double*)ccw_smem + warp_off;
    double* xo = xn + CCW_XW;
    double* rr = xo + CCW_XW;
    const PitchCfg& g = p.cfg[c.cls[clip]];
    const int W = g.nsamp_window, np = g.nsamp_period;
    const long long base = c.off[clip];
    const int nx = (int)(c.off[clip + 1] - base);
    const int16_t* pcm = pcm_all + base;
    const double x1 = c.x1[clip], t1 = p.t1[clip], dx = c.dx;
    constexpr int NL = 32 * TL;                   // lags 1 .. NL are carried

    double C[TL];
    int A0i = 0, Mi = 0;
    long long Q0i = 0;
    int lo_prev = 0, left_prev = 0;
    bool have = false;
#pragma unroll 1
    for (int kk = 0; kk < nseg; kk++) {
        long long lo_ll, left_ll;
        int Lloc;
        ccs_geom(g, x1, dx, t1, k0 + kk, (long long)nx, lo_ll, Lloc, left_ll);
        const int lo = (int)lo_ll, left = (int)left_ll;
        const int d = lo - lo_prev, dl = left - left_prev;
        const bool sliding = have && d >= 0 && d <= CCW_DMAX && dl >= 0 && dl <= CCW_DMAX;
        int j0 = 0, dn = 0;                       // afterwards xn[dn + e] = s[hi + e], xo[dn + e] = s[lo + e]
        if (!sliding) {                           // the window is slid open from empty, in blocks of <= CCW_DMAX samples
#pragma unroll
            for (int u = 0; u < TL; u++) C[u] = 0.0;
            A0i = 0; Q0i = 0; Mi = 0;
        }
#pragma unroll 1
        do {
            int nb, ob, len, olen;
            if (sliding) { nb = lo_prev + W; ob = lo_prev; len = d; olen = d; }
            else { len = W - j0 < CCW_DMAX ? W - j0 : CCW_DMAX; nb = lo + j0; ob = lo - len; olen = 0; j0 += len; }
            int sN, sO;
            long long qN, qO;
            __syncwarp();
            ccw_load(pcm, nx, nb, len + NL + 1, len, warp_off, sN, qN);
            ccw_load(pcm, nx, ob, len + NL + 1, olen, warp_off + CCW_XW, sO, qO);
            A0i += sN - sO;
            Q0i += qN - qO;
            __syncwarp();
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                // C + new - old = -(-(C + new) + old): the subtraction costs two sign flips instead of a multiply per sample
                if (pass == 1) {
#pragma unroll
                    for (int u = 0; u < TL; u++) C[u] = -C[u];
                }
                ccw_products<TL>(pass ? xo : xn, pass ? olen : len, lane, C);
                if (pass == 1) {
#pragma unroll
                    for (int u = 0; u < TL; u++) C[u] = -C[u];
                }
            }
            dn = len;
        } while (!sliding && j0 < W);
        if (sliding) Mi += ccw_isum2(pcm, nx, left_prev + np + 1, left + np, left_prev - np + 1, left - np);
        else Mi = ccw_isum2(pcm, nx, left - np + 1, left + np, 1, 0);
        have = true;
        lo_prev = lo; left_prev = left;

        // ---- raw-sample sums of every lag (exact), handed to the rolled normalisation loop through shared memory
        const double mu = ((double)Mi * (1.0 / 32768.0)) / (double)(2 * np);
        const double A0 = (double)A0i * (1.0 / 32768.0), Q0 = (double)Q0i * (1.0 / 1073741824.0);
        double da[TL], dq[TL];
        {
            const double* pn = xn + dn + TL * lane;
            const double* po = xo + dn + TL * lane;
            double sa = 0.0, sq = 0.0;
#pragma unroll
            for (int u = 0; u < TL; u++) {
                const double a = pn[u], b = po[u];
                sa += a - b;
                sq += fma(a, a, -(b * b));
                da[u] = sa; dq[u] = sq;
            }
            double ia = sa, iq = sq;                              // inclusive warp scan of the lane totals
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double ua = __shfl_up_sync(FULL_MASK, ia, o), uq = __shfl_up_sync(FULL_MASK, iq, o);
                if (lane >= o) { ia += ua; iq += uq; }
            }
            const double ea = ia - sa, eq = iq - sq;              // exclusive: sum over the lower lanes
            __syncwarp();                                         // every lane has read its window entries
#pragma unroll
            for (int u = 0; u < TL; u++) {
                const int l = 1 + TL * lane + u;
                xn[l] = C[u];
                xo[l] = A0 + (da[u] + ea);
                rr[l] = Q0 + (dq[u] + eq);
            }
        }
        __syncwarp();
        ccw_finish(c, p, g, pcm, nx, clip, f + kk, left, Lloc, NL, mu, A0, Q0, warp_off);
    }
}

template <int OCC>
__global__ void __launch_bounds__(CCW_NT, OCC) k_cc_frames_w(const __grid_constant__ Clips c, const __grid_constant__ PitchPass p,
                                                           const __grid_constant__ CcwParams A) {
    const int lane = threadIdx.x & 31;
    const int total = p.fstart[c.n];
    const int nturn = (total + A.run - 1) / A.run;
    for (;;) {
        int turn = 0;
        if (lane == 0) turn = atomicAdd(p.turn_counter, 1);
        turn = __shfl_sync(FULL_MASK, turn, 0);
        if (turn >= nturn) break;
        int f = turn * A.run;
        const int f_end = f + A.run < total ? f + A.run : total;
        while (f < f_end) {
            const int clip = find_segment(p.fstart, c.n, f);
            const int cend = p.fstart[clip + 1];
            const int nseg = (f_end < cend ? f_end : cend) - f;
            const int k0 = f - p.fstart[clip];
            const int Lm = p.cfg[c.cls[clip]].maximumLag;
            if (Lm + 1 <= 32 * 5) ccw_segment<5>(c, p, A.pcm, clip, f, nseg, k0);
            else if (Lm + 1 <= 32 * 7) ccw_segment<7>(c, p, A.pcm, clip, f, nseg, k0);
            else ccw_segment<9>(c, p, A.pcm, clip, f, nseg, k0);
            __syncwarp();
            f += nseg;
        }
    }
}

// returns false when the pass cannot run on the warp kernel (float64 input, more than 288 lags, rows that do not fit)
bool launch_cc_frames_warp(const Clips& c, const PitchPass& p, int max_frames_hint, cudaStream_t s) {
    if (c.pcm.p64 || !c.pcm.p16) return false;
    for (int k = 0; k < 3; k++) {
        const PitchCfg& g = p.cfg[k];
        if (g.method < 2) return false;
        if (g.maximumLag + 1 > 32 * 9 || g.maximumLag + 2 > CCW_RR || g.nsamp_window < 1) return false;
        if (g.maximumLag > g.brent_ixmax) return false;
    }
    CcwParams A;
    A.pcm = c.pcm.p16;
    const size_t smem = (size_t)CCW_WARPS * CCW_WARP_DOUBLES * sizeof(double);
    static int occ_want = 0;
    if (!occ_want) { const char* e = getenv("MSHDS_CCW_OCC"); occ_want = e && atoi(e) == 3 ? 3 : 2; }     // development switch (A/B)
    const void* kfn = occ_want == 3 ? (const void*)k_cc_frames_w<3> : (const void*)k_cc_frames_w<2>;
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, CCW_NT, smem);
    if (occ < 1) occ = 1;
    const int warps = sm_count() * occ * CCW_WARPS;
    // a run pays ~7 frames of start-up: long runs, but several turns per warp so that the tail stays short
    int run = max_frames_hint > 0 ? (max_frames_hint + 4 * warps - 1) / (4 * warps) : 128;
    if (run < 32) run = 32;
    if (run > 96) run = 96;
    static int run_env = -1;
    if (run_env < 0) { const char* e = getenv("MSHDS_CCW_RUN"); run_env = e ? atoi(e) : 0; }      // development switch (A/B)
    if (run_env > 0) run = run_env;
    A.run = run;
    int grid = sm_count() * occ;
    const int nturn = (max_frames_hint + run - 1) / run;
    const int need = (nturn + CCW_WARPS - 1) / CCW_WARPS;
    if (max_frames_hint > 0 && grid > need) grid = need;
    if (grid < 1) grid = 1;
    cudaMemsetAsync(p.turn_counter, 0, sizeof(int), s);
    if (p.hnr_mode) cudaMemsetAsync(p.qcount64, 0, sizeof(unsigned long long), s);
    else cudaMemsetAsync(p.qcount, 0, sizeof(int), s);
    if (occ_want == 3) k_cc_frames_w<3><<<grid, CCW_NT, smem, s>>>(c, p, A);
    else k_cc_frames_w<2><<<grid, CCW_NT, smem, s>>>(c, p, A);
    return true;
}

// returns false when the pass has to stay on the frame-by-frame kernel (float64 input, a frame step that is not a whole
// number of samples, scratch that does not fit)
bool launch_cc_frames_shared(const Clips& c, const PitchPass& p, int max_frames_hint, cudaStream_t s) {
    if (c.pcm.p64 || !c.pcm.p16) return false;
    int H = 0, Wmax = 0, Lmax = 0;
    for (int k = 0; k < 3; k++) {
        const PitchCfg& g = p.cfg[k];
        if (g.method < 2) return false;
        const double h = g.dt / c.dx;
        const int hk = (int)floor(h + 0.5);
        if (hk < 16 || fabs(h - hk) > 1e-9 * h) return false;
        if (H && hk != H) return false;
        H = hk;
        if (g.nsamp_window > Wmax) Wmax = g.nsamp_window;
        if (g.maximumLag > Lmax) Lmax = g.maximumLag;
    }
    if (H + Lmax + 2 * CCS_TL + 2 > CCS_MIRROR) return false;
    CcsParams A;
    A.pcm = c.pcm.p16;
    A.H = H;
    A.R = (Wmax + H - 1) / H + 3;
    int cap = 256;
    while (cap < Wmax + Lmax + 4 * H + 16) cap <<= 1;
    A.cap = cap;
    A.LS = ((Lmax + CCS_TL - 1) / CCS_TL) * CCS_TL + 1;
    A.nchunk_max = 8;
    int o = 0;
    auto take = [&](size_t bytes) { int r = o; o = (int)((o + bytes + 15) & ~(size_t)15); return r; };
    A.o_x = take(sizeof(double) * (cap + CCS_MIRROR));
    A.o_ps = take(sizeof(double) * cap);
    A.o_pq = take(sizeof(double) * cap);
    A.o_ring = take(sizeof(double) * (size_t)A.R * A.LS);
    A.o_part = take(sizeof(double) * (size_t)A.nchunk_max * A.LS);
    A.o_rr = take(sizeof(double) * (Lmax + 8));
    const size_t smem = (size_t)o;
    if (smem > 112 * 1024) return false;
    cudaFuncSetAttribute(k_cc_frames_s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_cc_frames_s, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_cc_frames_s, CCS_NT, smem);
    if (occ < 1) occ = 1;
    const int slots = sm_count() * occ;
    int run = max_frames_hint > 0 ? (max_frames_hint + slots - 1) / slots : 160;
    static int run_env = -1;
    if (run_env < 0) { const char* e = getenv("MSHDS_CCS_RUN"); run_env = e ? atoi(e) : 0; }      // development switch (A/B)
    if (run < 32) run = 32;
    if (run > 160) run = 160;
    if (run_env > 0) run = run_env;
    A.run = run;
    int grid = slots;
    const int nturn = (max_frames_hint + run - 1) / run;
    if (max_frames_hint > 0 && grid > nturn) grid = nturn;
    if (grid < 1) grid = 1;
    cudaMemsetAsync(p.turn_counter, 0, sizeof(int), s);
    if (p.hnr_mode) cudaMemsetAsync(p.qcount64, 0, sizeof(unsigned long long), s);
    else cudaMemsetAsync(p.qcount, 0, sizeof(int), s);
    k_cc_frames_s<<<grid, CCS_NT, smem, s>>>(c, p, A);
    return true;
}
