// k_pitch.cu -- Boersma (1993) pitch analysis on the GPU: autocorrelation (FFT) and forward cross-correlation frames,
// candidate picking with sinc refinement, Viterbi path finder, and the per-clip pitch statistics.
//
// Replaces the Praat calls behind mshds_extractor.py:104 (speech-rate pitch), :143 (_pitch_values), :178/:355
// (to_pitch_ac), :221 (to_harmonicity_cc), :241 (pitch inside "To Ltas (pitch-corrected)"), :270 (CPP pitch),
// :320 (to_pitch_cc) -- i.e. fon/Sound_to_Pitch.cpp Sound_to_Pitch_any + Sound_into_PitchFrame and
// fon/Pitch.cpp Pitch_pathFinder.
//
// Frame kernels.  The defaults live in their own files: autocorrelation frames with transforms of <= 1024 complex points run
// warp-per-frame on register-resident FFTs fed by TMA-staged spans (k_acw.cu), int16 cross-correlation frames as exact sliding
// sums, one warp per run of frames (k_ccs.cu).  The kernel HERE, k_pitch_frames, is the CTA-per-frame version of round 1
// (128 or 256 threads per frame, persistent grid over the flattened frame list, 8 consecutive frames per turn, the frame staged in
// shared memory, packed real FFT of fft.cuh for AC or register-tiled lag windows for FCC, first pass -- maxima, parabolic
// frequency, sinc30 strength, Praat's slot rule -- in the CTA).  It still serves the 4096-point speech-rate pass (:104), float64
// input behind the resampling front-end (cross-correlation) and the development switches "legacy_fft" / "legacy_cc" = 1, under
// which the parity tests compare it with the new kernels.  In every variant the correlation row and the <= 15 candidates leave
// the SM; the candidates that can matter are queued for the refinement kernels (sinc70/700 + Brent, 4 / 8 lanes per item), then
// k_pitch_score and the CTA-per-clip Viterbi (issued on a side stream by mshds_api.cu).
#include <cstdlib>
#include "internal.h"
#include "common.cuh"
#include "fft.cuh"
#include "num.cuh"

#define MAXPK 320
#define FRAMES_PER_TURN 8
#ifndef CC_TL_LONG
#define CC_TL_LONG 13
#endif

// ------------------------------------------------------------------------------------------------ frame grid
__global__ void k_pitch_grid(Clips c, PitchPass p) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    long long nx = c.off[i + 1] - c.off[i];
    const PitchCfg& g = p.cfg[c.cls[i]];
    int nf = 0;
    double t1 = 0.0;
    double duration = c.dx * (double)nx;
    bool ok = nx > 0 && !(g.floor_hz < g.ppw / duration) && g.halfnsamp_window >= 2;
    if (ok) ok = short_term_analysis(nx, c.dx, c.x1[i], g.grid_window, g.dt, &nf, &t1) != 0;
    if (!ok) nf = 0;
    p.nF[i] = nf;
    p.t1[i] = t1;
}

void launch_pitch_grid(const Clips& c, const PitchPass& p, cudaStream_t s) {
    k_pitch_grid<<<(c.n + 127) / 128, 128, 0, s>>>(c, p);
    launch_exclusive_scan(p.nF, p.fstart, c.n, s);
}

// ------------------------------------------------------------------------------------------------ candidates
struct CandScratch {
    double* rs0;        // [2Bs+1] symmetric correlation, rs0[Bs+i] = r[i], Bs = min(B, maximumLag + 32): the first pass
                        // never looks further than a depth-30 window around a lag below maximumLag
    double* pk_f;       // [MAXPK]
    double* pk_s;
    double* pk_key;
    int* pk_lag;
    unsigned* masks;
    double* cf;         // [MAXCAND+1] 1-based candidate slots
    double* cs;
    double* ckey;
    int* cimax;
    int* s_int;         // [4]: n maxima, ncand
    int pkcap;
};

// Praat's slot assignment (first free slot, else replace the weakest when stronger), run by ONE thread over the ordered
// list of maxima; only maxima with r > thr take part.  Returns ncand (>= 1, slot 1 = voiceless).
__device__ __forceinline__ int insert_candidates(const PitchCfg& g, const CandScratch& S, const double* r, int nmax, double thr,
                                                 double* cf, double* cs, double* ckey, int* cimax) {
    int ncand = 1;
    cf[1] = 0.0; cs[1] = 0.0; cimax[1] = 0;
    for (int m = 0; m < nmax; m++) {
        if (!(r[S.pk_lag[m]] > thr)) continue;
        int place = 0;
        if (ncand < g.maxn) {
            place = ++ncand;
        } else {
            double weakest = 2;
            for (int iweak = 2; iweak <= g.maxn; iweak++) {
                double ls = ckey[iweak];
                if (ls < weakest) { weakest = ls; place = iweak; }
            }
            if (S.pk_key[m] <= weakest) place = 0;
        }
        if (place) { cf[place] = S.pk_f[m]; cs[place] = S.pk_s[m]; ckey[place] = S.pk_key[m]; cimax[place] = S.pk_lag[m]; }
    }
    return ncand;
}

// Sound_into_PitchFrame, first pass: local maxima of r -> parabolic frequency + sinc(30) strength -> candidate slots.
// `vt2 >= 0` additionally builds the slots of a second analysis that differs only in its voicing threshold (the maxima
// of the lower threshold are a superset, the first-pass values are shared).  Returns ncand | ncand2 << 8.
// Harmonicity pass (maxn = 133 never fills, all path costs zero): only the list of maxima is built; returns their number.
template <int NT>
__device__ __forceinline__ int find_candidates(const PitchCfg& g, double dx, const CandScratch& S, int B, int Bs, int hnr_mode,
                                               const double2* __restrict__ tw, double vt2, double* cf2, double* cs2,
                                               double* ckey2, int* cimax2) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double thr1 = 0.5 * g.vt;
    const double thr = (vt2 >= 0.0 && 0.5 * vt2 < thr1) ? 0.5 * vt2 : thr1;
    int upper = g.maximumLag < B ? g.maximumLag : B;      // i < maximumLag && i < brent_ixmax
    int nlag = upper - 2;                                   // lags 2 .. upper-1
    if (nlag < 0) nlag = 0;
    int nrounds = (nlag + NT - 1) / NT;
    const double* r = S.rs0 + Bs;                           // r[i], i in [-Bs, Bs]
    for (int round = 0; round < nrounds; round++) {
        int i = 2 + round * NT + tid;
        bool flag = false;
        if (i < 2 + nlag) {
            double ri = r[i];
            flag = ri > thr && ri > r[i - 1] && ri >= r[i + 1];
        }
        unsigned m = __ballot_sync(FULL_MASK, flag);
        if (lane == 0) S.masks[round * (NT / 32) + warp] = m;
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int w = 0; w < nrounds * (NT / 32); w++) {
            unsigned m = S.masks[w];
            while (m) {
                int b = __ffs(m) - 1;
                m &= m - 1;
                if (n < S.pkcap) S.pk_lag[n++] = 2 + w * 32 + b;
            }
        }
        S.s_int[0] = n;
    }
    __syncthreads();
    const int nmax = S.s_int[0];
    if (hnr_mode) return nmax;          // harmonicity: every maximum is refined, the caller queues them
    const double* y1 = S.rs0 - 1;                           // 1-based view: y1[j] = r[j - B - 1]
    const int ny = 2 * Bs + 1;
    for (int m = warp; m < nmax; m += (NT / 32)) {
        int i = S.pk_lag[m];
        double dr = 0.5 * (r[i + 1] - r[i - 1]), d2r = 2 * r[i] - r[i - 1] - r[i + 1];
        double freq = 1.0 / dx / (i + dr / d2r);
        double x = 1.0 / dx / freq + (double)(Bs + 1);
        double strength = sinc_interp_warp(y1, ny, x, 30, lane, tw);
        if (strength > 1.0) strength = 1.0 / strength;
        if (lane == 0) {
            S.pk_f[m] = freq;
            S.pk_s[m] = strength;
            S.pk_key[m] = strength - g.octave_cost * log2(g.floor_hz / freq);
        }
    }
    __syncthreads();
    if (tid == 0) S.s_int[1] = insert_candidates(g, S, r, nmax, thr1, S.cf, S.cs, S.ckey, S.cimax);
    if (tid == 32 && vt2 >= 0.0) S.s_int[2] = insert_candidates(g, S, r, nmax, 0.5 * vt2, cf2, cs2, ckey2, cimax2);
    __syncthreads();
    return S.s_int[1] | ((vt2 >= 0.0 ? S.s_int[2] : 0) << 8);
}

struct FrameInfo {
    int clip, k;            // clip index, 0-based frame index in the clip
    long long base;         // sample offset of the clip
    long long nx;
    int cls;
};

// ------------------------------------------------------------------------------------------------ frame kernel
// IS_CC = false: autocorrelation (AC_HANNING);  true: forward cross-correlation (FCC_NORMAL), optionally HNR mode.
struct FrameSmem {      // byte offsets into dynamic shared memory (computed on the host)
    int a, rs, pkf, pks, pkkey, cf, cs, ckey, cf2, cs2, ckey2, cimax2, red, part, pklag, cimax, masks, sint, fi, total;
    int pkcap;          // capacity of the list of maxima
    int part_stride;    // CC: doubles per partial-sum row (>= maximumLag)
    int nchunk_max;
};


// Forward cross-correlation products sum_j x[j] * x[j + lag] for lag = 1..Lmax over the window j < W.  A thread owns TL
// consecutive lags and a chunk of the window: per step one new sample enters a sliding register window and feeds TL FMAs
// (2 shared-memory loads per TL float64 FMAs).  Partial sums go to part[chunk][lag-1]; returns the number of chunks.
template <int TL, int NT>
__device__ __forceinline__ int cc_products(const double* xs, double* part, int PS, int W, int Lmax, int nchunk_max) {
    const int tid = threadIdx.x;
    const int ngroups = (Lmax + TL - 1) / TL;
    int nchunk = ngroups > 0 ? NT / ngroups : 1;
    if (nchunk < 1) nchunk = 1;
    if (nchunk > nchunk_max) nchunk = nchunk_max;
    const int q = (W + nchunk - 1) / nchunk;
    for (int wi = tid; wi < ngroups * nchunk; wi += NT) {
        const int grp = wi % ngroups, ch = wi / ngroups;
        const int lag0 = 1 + TL * grp;
        const int j0 = ch * q, j1 = j0 + q < W ? j0 + q : W;
        double acc[TL], yw[TL];
        const double* xa = xs + j0;
        const double* xb = xs + j0 + lag0;
#pragma unroll
        for (int u = 0; u < TL; u++) { acc[u] = 0.0; yw[u] = xb[u]; }
        const int len = j1 - j0;
        int j = 0;
        // TL steps per trip with the window kept as a ring (sample xb[k] sits in slot k % TL): no register shuffling, the
        // trip is 2*TL shared-memory loads and TL*TL FMAs
        for (; j + TL <= len; j += TL) {
#pragma unroll
            for (int t = 0; t < TL; t++) {
                const double xv = xa[j + t];
#pragma unroll
                for (int u = 0; u < TL; u++) acc[u] = fma(xv, yw[(u + t) % TL], acc[u]);
                yw[t] = xb[j + t + TL];
            }
        }
        for (; j < len; j++) {          // tail (< TL steps): slots are aligned again, shift the window
            const double xv = xa[j];
#pragma unroll
            for (int u = 0; u < TL; u++) acc[u] = fma(xv, yw[u], acc[u]);
#pragma unroll
            for (int u = 0; u < TL - 1; u++) yw[u] = yw[u + 1];
            yw[TL - 1] = xb[j + TL];
        }
        double* pr = part + (size_t)ch * PS + (lag0 - 1);
#pragma unroll
        for (int u = 0; u < TL; u++) pr[u] = acc[u];
    }
    return nchunk;
}

template <bool IS_CC, int NT>
__global__ void __launch_bounds__(NT, (IS_CC ? 512 : 1024) / NT) k_pitch_frames(Clips c, PitchPass p, const double2* __restrict__ tw, FrameSmem L) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* a = (double2*)(smem + L.a);            // AC: packed FFT buffer; CC: xs[] doubles
    double* xs = (double*)(smem + L.a);
    CandScratch S;
    S.rs0 = (double*)(smem + L.rs);
    S.pk_f = (double*)(smem + L.pkf);
    S.pk_s = (double*)(smem + L.pks);
    S.pk_key = (double*)(smem + L.pkkey);
    S.cf = (double*)(smem + L.cf);
    S.cs = (double*)(smem + L.cs);
    S.ckey = (double*)(smem + L.ckey);
    double* cf2 = (double*)(smem + L.cf2);
    double* cs2 = (double*)(smem + L.cs2);
    double* ckey2 = (double*)(smem + L.ckey2);
    int* cimax2 = (int*)(smem + L.cimax2);
    double* red = (double*)(smem + L.red);
    double* part = (double*)(smem + L.part);
    S.pk_lag = (int*)(smem + L.pklag);
    S.cimax = (int*)(smem + L.cimax);
    S.masks = (unsigned*)(smem + L.masks);
    S.s_int = (int*)(smem + L.sint);
    S.pkcap = L.pkcap;
    FrameInfo* fi = (FrameInfo*)(smem + L.fi);
    const int PS = L.part_stride;

    const int tid = threadIdx.x;
    const int total = p.fstart[c.n];
    const double dx = c.dx;

    // a CTA takes FRAMES_PER_TURN consecutive frames per turn: neighbouring frames share ~90 % of their samples (L1 hits)
    // turns are handed out by an atomic counter: frames of low-pitched speakers cost up to twice those of high-pitched ones,
    // and a fixed round-robin share leaves the last CTAs running alone at the end of the launch
    const int nturn = (total + FRAMES_PER_TURN - 1) / FRAMES_PER_TURN;
    __shared__ int s_turn;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_turn = atomicAdd(p.turn_counter, 1);
        __syncthreads();
        const int turn = s_turn;
        if (turn >= nturn) break;
    for (int f = turn * FRAMES_PER_TURN; f < total && f < (turn + 1) * FRAMES_PER_TURN; f++) {
        __syncthreads();
        if (tid == 0) {
            int clip = find_segment(p.fstart, c.n, f);
            fi->clip = clip;
            fi->k = f - p.fstart[clip];
            fi->base = c.off[clip];
            fi->nx = c.off[clip + 1] - c.off[clip];
            fi->cls = c.cls[clip];
        }
        __syncthreads();
        const int clip = fi->clip;
        const double x1 = c.x1[clip];
        const PitchCfg& g = p.cfg[fi->cls];
        const SPtr pcm = c.pcm + fi->base;                // pcm[i-1] is 1-based sample i
        const long long nx = fi->nx;
        const double t = p.t1[clip] + (double)fi->k * g.dt;
        const long long leftSample = x_to_low(x1, dx, t), rightSample = leftSample + 1;
        const double globalPeak = c.gpeak[clip];
        const int B = g.brent_ixmax;
        const int Bs = B < g.maximumLag + 32 ? B : g.maximumLag + 32;      // extent of the shared-memory copy (see CandScratch)
        const int W = g.nsamp_window;

        // local mean over one longest period to both sides
        double acc = 0.0;
        {
            long long s0 = rightSample - g.nsamp_period, s1 = leftSample + g.nsamp_period;
            for (long long i = s0 + tid; i <= s1; i += NT) acc += samp(pcm, i - 1);
        }
        const double localMean = block_sum(acc, red) / (double)(2 * g.nsamp_period);

        // frame copy (+ window for AC) and local peak
        double lp = 0.0;
        const long long startSample = rightSample - g.halfnsamp_window;
        int pk0 = g.halfnsamp_window + 1 - g.halfnsamp_period; if (pk0 < 1) pk0 = 1;
        int pk1 = g.halfnsamp_window + g.halfnsamp_period; if (pk1 > W) pk1 = W;
        if (!IS_CC) {
            double* ar = (double*)a;                         // packed: ar[m] = frame[m+1]
            const int Nfft = g.nsampFFT;
            for (int m = tid; m < Nfft; m += NT) {
                double v = 0.0;
                if (m < W) {
                    v = (samp(pcm, startSample + m - 1) - localMean) * __ldg(g.window + m);
                    int j = m + 1;
                    if (j >= pk0 && j <= pk1) lp = fmax(lp, fabs(v));
                }
                ar[SWZD(m)] = v;
            }
        } else {
            for (int m = tid; m < W; m += NT) {
                int j = m + 1;
                if (j >= pk0 && j <= pk1) lp = fmax(lp, fabs(samp(pcm, startSample + m - 1) - localMean));
            }
        }
        const double localPeak = block_max(lp, red);
        const double intensity = localPeak > globalPeak ? 1.0 : localPeak / globalPeak;
        __syncthreads();

        const int Ls = stored_lags(g);
        double* rrow = p.rbuf + (size_t)f * p.rstride;
        if (!IS_CC) {
            fft_dif<-1>(a, g.M, tw);
            packed_power_to_inverse_input(a, g.M, g.logM, tw, IdentityF(), (double*)nullptr);
            fft_dit<+1>(a, g.M, tw);
            const double* ac = (const double*)a;            // ac[i] natural order
            const double ac0 = ac[SWZD(0)];
            for (int i = tid; i <= B; i += NT) {
                double v = i == 0 ? 1.0 : ac[SWZD(i)] / (ac0 * __ldg(g.windowR + i));
                if (i <= Bs) { S.rs0[Bs + i] = v; S.rs0[Bs - i] = v; }
                rrow[i] = v;
            }
        } else {
            // forward cross-correlation
            const double startTime = t - 0.5 * (1.0 / g.floor_hz + g.dt_window);
            long long startS = x_to_low(x1, dx, startTime);
            if (startS < 1) startS = 1;
            long long localSpan = g.maximumLag + W;
            if (localSpan > nx + 1 - startS) localSpan = nx + 1 - startS;
            const int localMaximumLag = (int)(localSpan - W);
            const int Lmax = localMaximumLag > 0 ? localMaximumLag : 0;
            // xs[j-1] = s[startS-1+j] - localMean, j = 1..localSpan; zero tail so the tiled loop may read ahead
            const int xs_len = g.maximumLag + W + 32;      // read-ahead of the tiled loop: < 2 * TL + 1 samples past the span
            for (int j = tid; j < xs_len; j += NT) xs[j] = j < (int)localSpan ? samp(pcm, startS - 1 + j) - localMean : 0.0;
            for (int i = tid; i < 2 * Bs + 1; i += NT) S.rs0[i] = 0.0;
            for (int i = tid; i < Ls; i += NT) rrow[i] = 0.0;
            __syncthreads();
            // prefix sums of squares: sq[k] = sum_{j<k} xs[j]^2  (stored behind the partial products)
            double* sq = part + (size_t)L.nchunk_max * PS;
            {
                const int total_len = (int)localSpan;
                const int per = (total_len + NT - 1) / NT;
                const int b0 = tid * per, b1 = b0 + per < total_len ? b0 + per : total_len;
                double loc = 0.0;
                for (int j = b0; j < b1; j++) loc = fma(xs[j], xs[j], loc);
                // exclusive scan of the per-thread sums
                const int lane = tid & 31, w = tid >> 5;
                double incl = loc;
                for (int o = 1; o < 32; o <<= 1) { double y = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += y; }
                if (lane == 31) red[w] = incl;
                __syncthreads();
                double woff = 0.0;
                for (int k = 0; k < w; k++) woff += red[k];
                double run = woff + incl - loc;
                for (int j = b0; j < b1; j++) { sq[j] = run; run = fma(xs[j], xs[j], run); }
                if (b1 == total_len && b0 < total_len) sq[total_len] = run;
                if (total_len == 0 && tid == 0) sq[0] = 0.0;
                __syncthreads();
            }
            const double sumx2 = sq[W] - sq[0];
            // products: work item = (group of TL lags, chunk of the window).  TL is odd so that the lag windows of adjacent
            // lanes start TL doubles apart (an even stride would put the whole warp on 2-4 shared-memory banks)
            int nchunk;
            // (shared memory, not registers, limits the CTAs per SM here, so the long-window pass can afford 13 lags per thread:
            // 2 loads per 13 FMAs)
            if (W >= 600) nchunk = cc_products<CC_TL_LONG, NT>(xs, part, PS, W, Lmax, L.nchunk_max);
            else nchunk = cc_products<5, NT>(xs, part, PS, W, Lmax, L.nchunk_max);
            __syncthreads();
            for (int lag = 1 + tid; lag <= Lmax; lag += NT) {
                double pr = 0.0;
                for (int ch = 0; ch < nchunk; ch++) pr += part[(size_t)ch * PS + lag - 1];
                double sy = sq[lag + W] - sq[lag];
                double v = pr / sqrt(sumx2 * sy);
                S.rs0[Bs + lag] = v;
                S.rs0[Bs - lag] = v;
                if (lag < Ls) rrow[lag] = v;
            }
            if (tid == 0) { S.rs0[Bs] = 1.0; rrow[0] = 1.0; }
        }
        __syncthreads();

        if (p.hnr_mode) {
            // Sound_to_Harmonicity_cc: the frame value is the best refined strength over ALL maxima of r; queue every one
            // unvoiced local score: a voiced candidate needs strength > 2 - intensity/sil (vt = 0) and strengths never
            // exceed 1 (reflection), so a frame with 2 - intensity/sil >= 1 is -200 dB whatever its candidates are
            double uvs = g.sil <= 0 ? 0.0 : 2.0 - intensity / (g.sil / (1.0 + g.vt));
            uvs = g.vt + (uvs > 0 ? uvs : 0);
            int nmax = 0;
            if (localPeak != 0.0 && uvs < 1.0) nmax = find_candidates<NT>(g, dx, S, B, Bs, 1, tw, -1.0, cf2, cs2, ckey2, cimax2);
            const double* r = S.rs0 + Bs;
            for (int m = tid; m < nmax; m += NT) {
                const int i = S.pk_lag[m];
                double dr = 0.5 * (r[i + 1] - r[i - 1]), d2r = 2 * r[i] - r[i - 1] - r[i + 1];
                double freq = 1.0 / dx / (i + dr / d2r);
                unsigned long long item = ((unsigned long long)(unsigned)f << 32) | ((unsigned long long)i << 8) |
                                          (freq > 0.3 / dx ? 1ull : 0ull);
                unsigned long long slot = atomicAdd(p.qcount64, 1ull);
                if (slot < p.q64_cap) p.queue64[slot] = item;
                else atomicOr(&c.status[clip], ST_HNR);              // cannot happen: capacity is the worst case
            }
            if (tid == 0) { p.inten[f] = intensity; p.best_bits[f] = 0ull; }
            continue;
        }
        const bool dual = p.dual_cand_f != nullptr;
        int ncand = 1, ncand2 = 1;
        if (localPeak != 0.0) {
            int nn = find_candidates<NT>(g, dx, S, B, Bs, 0, tw, dual ? p.dual_vt : -1.0, cf2, cs2, ckey2, cimax2);
            ncand = nn & 0xff;
            if (dual) ncand2 = nn >> 8;
        } else if (tid == 0) {
            S.cf[1] = 0.0; S.cs[1] = 0.0; S.cimax[1] = 0;
            cf2[1] = 0.0; cs2[1] = 0.0; cimax2[1] = 0;
        }
        __syncthreads();

        // Pitch_pathFinder: replacing a voiced candidate of frame t by the voiceless one gains at least
        // (unvoicedStrength - 1) locally (voiced local scores are <= 1) and costs at most two voiced/unvoiced transitions,
        // so when unvoicedStrength > 1 + 2*vuvCost no voiced candidate of this frame can lie on the best path: their
        // refined values cannot matter and the refinement is skipped (exact).
        // Candidates leave the SM; the ones that can matter are queued for the refinement kernel.  Threads 0..14 serve
        // the analysis itself, threads 32..46 the second analysis that shares this correlation (dual mode).
        const int set = tid >> 5, ctid = tid & 31;
        if (ctid < MAXCAND && (set == 0 || (set == 1 && dual))) {
            const double vt = set == 0 ? g.vt : p.dual_vt;
            double uvs = g.sil <= 0 ? 0.0 : 2.0 - intensity / (g.sil / (1.0 + vt));
            uvs = vt + (uvs > 0 ? uvs : 0);
            const bool frame_stays_unvoiced = uvs > 1.0 + 2.0 * g.vuv_cost * (0.01 / g.dt) + 1e-9;
            const int nc = set == 0 ? ncand : ncand2;
            const double* scf = set == 0 ? S.cf : cf2;
            const double* scs = set == 0 ? S.cs : cs2;
            const int* sci = set == 0 ? S.cimax : cimax2;
            const int ci = ctid + 1;
            double fr = 0.0, st = 0.0;
            int im = 0;
            if (ci <= nc) { fr = scf[ci]; st = scs[ci]; im = sci[ci]; }
            const size_t o2 = (size_t)f * MAXCAND + ctid;
            if (set == 0) { p.cand_f[o2] = fr; p.cand_s[o2] = st; p.cand_imax[o2] = (unsigned short)im; }
            else { p.dual_cand_f[o2] = fr; p.dual_cand_s[o2] = st; p.dual_cand_imax[o2] = (unsigned short)im; }
            // A candidate whose refined lag cannot fall below fs/ceiling (refined lag <= imax+1) stays voiceless for the
            // path finder whatever its refined values: it is never refined.
            bool live = ci >= 2 && ci <= nc && (1.0 / dx / (double)(im + 1) < g.ceiling) && !frame_stays_unvoiced;
            if (live) {
                if (set == 0) { int slot = atomicAdd(p.qcount, 1); p.queue[slot] = f * 16 + ctid; }
                else { int slot = atomicAdd(p.dual_qcount, 1); p.dual_queue[slot] = f * 16 + ctid; }
            }
        }
        if (tid == 0) {
            p.ncand[f] = (uint8_t)ncand; p.inten[f] = intensity;
            if (dual) { p.dual_ncand[f] = (uint8_t)ncand2; p.dual_inten[f] = intensity; }
        }
    }
    }
}

static FrameSmem frames_smem_layout(const PitchPass& p, bool is_cc) {
    int ab = 0, rs = 0, ml = 0;
    for (int k = 0; k < 3; k++) {
        const PitchCfg& g = p.cfg[k];
        int need_a = is_cc ? (int)sizeof(double) * (g.maximumLag + g.nsamp_window + 40) : (int)sizeof(double2) * g.M;
        if (need_a > ab) ab = need_a;
        const int bs = g.brent_ixmax < g.maximumLag + 32 ? g.brent_ixmax : g.maximumLag + 32;
        if (2 * bs + 1 > rs) rs = 2 * bs + 1;
        if (g.maximumLag > ml) ml = g.maximumLag;
    }
    FrameSmem L;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o = (o + bytes + 15) & ~15; return r; };
    // the list of maxima (at most one every other lag) is built after the transform buffer / sample window is dead: it lives there
    int pkcap = ml / 2 + 2;
    if (pkcap > MAXPK) pkcap = MAXPK;
    L.pkcap = pkcap;
    const int pkbytes = ((int)sizeof(double) * pkcap + 15) & ~15;
    if (ab < 3 * pkbytes + (((int)sizeof(int) * pkcap + 15) & ~15)) ab = 3 * pkbytes + (((int)sizeof(int) * pkcap + 15) & ~15);
    L.a = take(ab);
    L.pkf = L.a; L.pks = L.a + pkbytes; L.pkkey = L.a + 2 * pkbytes; L.pklag = L.a + 3 * pkbytes;
    L.rs = take((int)sizeof(double) * (rs + 1));
    L.cf = take((int)sizeof(double) * (MAXCAND + 1));
    L.cs = take((int)sizeof(double) * (MAXCAND + 1));
    L.ckey = take((int)sizeof(double) * (MAXCAND + 1));
    L.cf2 = take((int)sizeof(double) * (MAXCAND + 1));
    L.cs2 = take((int)sizeof(double) * (MAXCAND + 1));
    L.ckey2 = take((int)sizeof(double) * (MAXCAND + 1));
    L.cimax2 = take((int)sizeof(int) * (MAXCAND + 1));
    L.red = take((int)sizeof(double) * 32);
    L.part_stride = (ml + 16 + 7) & ~7;
    L.nchunk_max = 8;
    // CC: nchunk_max rows of partial products + the prefix sums of squares (window + maximumLag + 1 entries)
    L.part = take(is_cc ? (int)sizeof(double) * (L.nchunk_max * L.part_stride + ab / (int)sizeof(double) + 8) : 16);
    L.cimax = take((int)sizeof(int) * (MAXCAND + 1));
    L.masks = take((int)sizeof(unsigned) * 64);
    L.sint = take((int)sizeof(int) * 4);
    L.fi = take(64);
    L.total = o;
    return L;
}

void launch_pitch_frames(const Clips& c, const PitchPass& p, const double2* tw, int max_frames_hint, cudaStream_t s) {
    bool is_cc = p.cfg[0].method != 0;
    // autocorrelation passes with transforms of <= 1024 complex points run warp-per-frame with TMA-staged samples (k_acw.cu)
    if (!is_cc && !c.legacy_fft && c.twb512 &&
        launch_ac_frames_warp(c, p, tw, c.twb512, c.twb1024, c.total_samples, max_frames_hint, s)) return;
    // int16 cross-correlation passes share their block products between overlapping frames (k_ccs.cu); the candidates of
    // to_pitch_cc then come from the warp-per-frame candidate kernel, the harmonicity pass queues its maxima itself
    if (is_cc && c.legacy_cc != 1 &&
        ((c.legacy_cc == 0 && launch_cc_frames_warp(c, p, max_frames_hint, s)) || launch_cc_frames_shared(c, p, max_frames_hint, s))) {
        if (!p.hnr_mode) launch_ac_candidates(c, p, tw, max_frames_hint, s);
        return;
    }
    FrameSmem L = frames_smem_layout(p, is_cc);
    size_t smem = (size_t)L.total;
    const int nsm = sm_count();
    int blocks_per_sm = (int)(220 * 1024 / (smem + 1024));
    if (blocks_per_sm > 6) blocks_per_sm = 6;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int grid = nsm * blocks_per_sm;
    if (max_frames_hint > 0 && grid > max_frames_hint) grid = max_frames_hint;
    if (grid < 1) grid = 1;
    cudaMemsetAsync(p.qcount, 0, sizeof(int), s);
    cudaMemsetAsync(p.turn_counter, 0, sizeof(int), s);
    if (p.dual_cand_f) cudaMemsetAsync(p.dual_qcount, 0, sizeof(int), s);
    if (p.hnr_mode) cudaMemsetAsync(p.qcount64, 0, sizeof(unsigned long long), s);
    static int nt_ac = -1, nt_cc = -1;
    if (nt_ac < 0) {   // development switches (128 / 256 threads per frame); 0 = the rule below
        const char* e = getenv("MSHDS_NT_AC");
        nt_ac = e && atoi(e) == 256 ? 256 : (e && atoi(e) == 128 ? 128 : 0);
        e = getenv("MSHDS_NT_CC");
        nt_cc = e && atoi(e) == 256 ? 256 : (e && atoi(e) == 128 ? 128 : 0);
    }
    // 128 threads per frame keep every thread busy in the radix-4 passes of the 512/1024-point transforms and make the ~20
    // barriers per frame cheaper (measured: wide AC pass 28.2 -> 25.5 ms, formant CC pass 17.3 -> 12.4 ms on 96 x 30 s);
    // the 2048-point transform of the 30 Hz speech-rate pass is better off with 256 (13.0 vs 15.0 ms)
    int maxM = 0;
    for (int k = 0; k < 3; k++) if (p.cfg[k].M > maxM) maxM = p.cfg[k].M;
    const int nt = is_cc ? (nt_cc ? nt_cc : 128) : (nt_ac ? nt_ac : (maxM >= 2048 ? 256 : 128));
    // persistent grid: exactly the CTAs that are resident at once (registers or shared memory, whichever binds)
#define PF_LAUNCH(CC, N) \
    do { \
        cudaFuncSetAttribute(k_pitch_frames<CC, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        cudaFuncSetAttribute(k_pitch_frames<CC, N>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
        int occ = 0; \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pitch_frames<CC, N>, N, smem); \
        if (occ < 1) occ = 1; \
        grid = nsm * occ; \
        const int nturn = (max_frames_hint + FRAMES_PER_TURN - 1) / FRAMES_PER_TURN; \
        if (max_frames_hint > 0 && grid > nturn) grid = nturn; \
        if (grid < 1) grid = 1; \
        k_pitch_frames<CC, N><<<grid, N, smem, s>>>(c, p, tw, L); \
    } while (0)
    if (is_cc) { if (nt == 128) PF_LAUNCH(true, 128); else PF_LAUNCH(true, 256); }
    else { if (nt == 128) PF_LAUNCH(false, 128); else PF_LAUNCH(false, 256); }
#undef PF_LAUNCH
}

// ------------------------------------------------------------------------------------------------ refinement
#define RF_HALF 71          // sinc70 evaluated on [c-1, c+1] touches y[c-70 .. c+70]
#define RF_WIN (2 * RF_HALF + 2)

// NUMimproveMaximum of candidate `imax` on the correlation row of frame f.  The 141 row values a sinc70 search can touch
// are staged once per item in shared memory (the search evaluates the interpolation ~13 times); sinc700 candidates
// (f > 0.3 fs, lag < 3.4) read the row directly.
// Lanes per refinement item (template parameter NL).  70 taps per side: 18 per lane on 4 lanes, 9 on 8 (97 % slot use; the
// per-evaluation table look-ups are amortised over the taps and the Brent bookkeeping replicated per lane is issued once per
// 8 / 4 items), 5 per lane on 16 (87 %).
template <int NL>
__device__ __forceinline__ double refine_candidate(const SymRowY& y, int B, int imax, bool deep, double* st, double* xmid_out,
                                                   int lane, unsigned mask, const double2* __restrict__ tw) {
    const int n = 2 * B + 1, c0 = imax + B + 1;
    double xmid, ymid;
    if (deep) {
        ymid = improve_extremum_warp_t(y, n, c0, PEAK_SINC700, &xmid, true, lane, tw, mask, NL);
    } else {
        const int j0 = c0 - RF_HALF;
        __syncwarp(mask);
        for (int j = lane; j < RF_WIN; j += NL) {
            int idx = j0 + j;
            st[j] = (idx >= 1 && idx <= n) ? y(idx) : 0.0;
        }
        __syncwarp(mask);
        StagedY ys{st, j0};
        ymid = improve_extremum_warp_t(ys, n, c0, PEAK_SINC70, &xmid, true, lane, tw, mask, NL);
    }
    *xmid_out = xmid - (double)(B + 1);
    return ymid;
}

// Sound_into_PitchFrame, second pass: NUMimproveMaximum with sinc(70/700) + Brent on the stored correlation row.  A flat,
// perfectly balanced work list: one lane group per queued (frame, candidate).
template <int NL>
__global__ void __launch_bounds__(256, 3) k_pitch_refine(Clips c, PitchPass p, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    double (*s_stage)[RF_WIN] = (double (*)[RF_WIN])rf_smem;
    const int lane = threadIdx.x & (NL - 1);
    const int grp = threadIdx.x / NL;
    const unsigned mask = NL == 32 ? FULL_MASK : (((1u << NL) - 1u) << ((threadIdx.x & 31) & ~(NL - 1)));
    double* st = s_stage[grp];
    const int gg = blockIdx.x * (blockDim.x / NL) + grp;
    const int ng = gridDim.x * (blockDim.x / NL);
    const int total = *p.qcount;
    const double dx = c.dx;
    for (int q = gg; q < total; q += ng) {
        const int item = p.queue[q];
        const int f = item >> 4, slot = item & 15;
        const int clip = find_segment(p.fstart, c.n, f);
        const PitchCfg& g = p.cfg[c.cls[clip]];
        const int B = g.brent_ixmax;
        SymRowY y;
        y.row = p.rbuf + (size_t)f * p.rstride;
        y.centre = B + 1;
        y.len = stored_lags(g);
        const size_t o2 = (size_t)f * MAXCAND + slot;
        const int imax = p.cand_imax[o2];
        const double f0 = p.cand_f[o2];
        double xmid;
        double ymid = refine_candidate<NL>(y, B, imax, f0 > 0.3 / dx, st, &xmid, lane, mask, tw);
        if (ymid > 1.0) ymid = 1.0 / ymid;
        if (lane == 0) { p.cand_f[o2] = 1.0 / dx / xmid; p.cand_s[o2] = ymid; }
    }
}

// Harmonicity variant: every maximum of every frame is an item (frame, lag, depth flag); the frame keeps the largest
// refined strength among candidates that stay below the Nyquist "ceiling" (atomicMax on the bits of a positive double).
template <int NL>
__global__ void __launch_bounds__(256, 3) k_hnr_refine(Clips c, PitchPass p, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char rf_smem[];
    double (*s_stage)[RF_WIN] = (double (*)[RF_WIN])rf_smem;
    const int lane = threadIdx.x & (NL - 1);
    const int grp = threadIdx.x / NL;
    const unsigned mask = NL == 32 ? FULL_MASK : (((1u << NL) - 1u) << ((threadIdx.x & 31) & ~(NL - 1)));
    double* st = s_stage[grp];
    const long long gg = (long long)blockIdx.x * (blockDim.x / NL) + grp;
    const long long ng = (long long)gridDim.x * (blockDim.x / NL);
    unsigned long long total = *p.qcount64;
    if (total > p.q64_cap) total = p.q64_cap;
    const double dx = c.dx;
    for (long long q = gg; q < (long long)total; q += ng) {
        const unsigned long long item = p.queue64[q];
        const int f = (int)(item >> 32), imax = (int)((item >> 8) & 0xffffff);
        const bool deep = (item & 1ull) != 0;
        const int clip = find_segment(p.fstart, c.n, f);
        const PitchCfg& g = p.cfg[c.cls[clip]];
        const int B = g.brent_ixmax;
        SymRowY y;
        y.row = p.rbuf + (size_t)f * p.rstride;
        y.centre = B + 1;
        y.len = stored_lags(g);
        double xmid;
        double ymid = refine_candidate<NL>(y, B, imax, deep, st, &xmid, lane, mask, tw);
        if (ymid > 1.0) ymid = 1.0 / ymid;
        const double fr = 1.0 / dx / xmid;
        if (lane == 0 && fr > 0.0 && fr < g.ceiling && ymid > 0.0)
            atomicMax(p.best_bits + f, (unsigned long long)__double_as_longlong(ymid));
    }
}

// (Round 2 tried to skip maxima that "cannot win" the frame -- refine the highest maximum first, drop every maximum with
// r[i] + d2r + 1e-3 below that strength.  The bench's own parity check caught it: on noise-like correlation rows the sinc
// interpolant overshoots the samples by far more than the second difference (tools measurement with the oracle: 371 of
// 13,024 maxima violate the bound, by up to 0.22), so no cheap bound is rigorous and every maximum stays refined.)
// Pitch_pathFinder local scores (Viterbi passes) / per-frame winner (harmonicity pass)
__global__ void k_pitch_score(Clips c, PitchPass p) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= p.fstart[c.n]) return;
    const int clip = find_segment(p.fstart, c.n, f);
    const PitchCfg& g = p.cfg[c.cls[clip]];
    const double intensity = p.inten[f];
    double unvoicedStrength = g.sil <= 0 ? 0.0 : 2.0 - intensity / (g.sil / (1.0 + g.vt));
    unvoicedStrength = g.vt + (unvoicedStrength > 0 ? unvoicedStrength : 0);
    if (p.hnr_mode) {
        // all path costs are zero: the path is the per-frame first maximum of the local scores (voiceless comes first)
        const unsigned long long bits = p.best_bits[f];
        const double st = __longlong_as_double((long long)bits);
        const bool voiced = bits != 0ull && st > unvoicedStrength;
        p.sel_f[f] = voiced ? 1.0 : 0.0;          // only the voiced flag of the winner is used downstream
        p.sel_s[f] = voiced ? st : DEVNAN;
        return;
    }
    const int ncand = p.ncand[f];
    for (int ci = 0; ci < MAXCAND; ci++) {
        double sc = -1e300, lf = -1.0;
        if (ci < ncand) {
            double fr = p.cand_f[(size_t)f * MAXCAND + ci], st = p.cand_s[(size_t)f * MAXCAND + ci];
            bool voiceless = !(fr > 0.0 && fr < g.ceiling);
            sc = voiceless ? unvoicedStrength : st - g.octave_cost * log2(g.ceiling / fr);
            lf = voiceless ? -1.0 : log2(fr);
        }
        p.cand_score[(size_t)f * MAXCAND + ci] = sc;
        p.cand_lf[(size_t)f * MAXCAND + ci] = lf;
    }
}

void launch_pitch_refine(const Clips& c, const PitchPass& p, const double2* tw, int max_frames_hint, cudaStream_t s) {
    (void)max_frames_hint;
    static int nl_env = -1;
    if (nl_env < 0) {   // development switch; the defaults are what the committed measurements use
        const char* e = getenv("MSHDS_RF_NL");
        nl_env = e ? atoi(e) : 0;
        if (nl_env != 4 && nl_env != 8 && nl_env != 16) nl_env = 0;
    }
    // measured on B200 (96 x 30 s): Viterbi passes 4 lanes 16.3 ms / 8 lanes 20.5 ms / 16 lanes 29.3 ms; harmonicity pass
    // (a third of its items are 700-tap ones) 34.4 / 31.9 / 39.5 ms
    const int nl = nl_env ? nl_env : (p.hnr_mode ? 8 : 4);
    const int grid = sm_count() * 3;
    const size_t smem = (size_t)(256 / nl) * RF_WIN * sizeof(double);
#define RF_LAUNCH(K, N) \
    do { \
        cudaFuncSetAttribute(K<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        K<N><<<grid, 256, smem, s>>>(c, p, tw); \
    } while (0)
    if (p.hnr_mode) {
        if (nl == 4) RF_LAUNCH(k_hnr_refine, 4); else if (nl == 8) RF_LAUNCH(k_hnr_refine, 8); else RF_LAUNCH(k_hnr_refine, 16);
    } else {
        if (nl == 4) RF_LAUNCH(k_pitch_refine, 4); else if (nl == 8) RF_LAUNCH(k_pitch_refine, 8); else RF_LAUNCH(k_pitch_refine, 16);
    }
#undef RF_LAUNCH
}
void launch_pitch_score(const Clips& c, const PitchPass& p, int max_frames_hint, cudaStream_t s) {
    int blocks = (max_frames_hint + 127) / 128;
    if (blocks < 1) blocks = 1;
    k_pitch_score<<<blocks, 128, 0, s>>>(c, p);
}

// ------------------------------------------------------------------------------------------------ Viterbi
// fon/Pitch.cpp Pitch_pathFinder.  The recurrence is evaluated with exactly the reference's arithmetic,
// value = (delta[c1] - cost) + score, and its tie rule (the first maximum wins), so the back-pointers are bit-identical to
// a sequential float64 loop.
//
// Latency matters here: a 60 s recording is a chain of 12,000 dependent steps and nothing else runs when the batch is one
// clip (BASELINE.json configs[0]).  Round 1 used one warp per clip with a serial compare chain over the <= 15 previous
// candidates: 1,080 cycles per frame; a warp-level rewrite with more instruction-level parallelism did not help -- ncu
// showed a single warp issuing one instruction every 5 cycles, i.e. the step is instruction-count bound (335 per frame).
// Now ONE CTA of 256 threads owns a clip: thread (c2, c1) = (tid / 16, tid % 16) forms the one value of its candidate
// pair, the 16 lanes of a half-warp reduce over c1 with four shuffle steps (ties go to the lower index), the state lives in
// double-buffered shared memory and each frame costs one block barrier.  Candidate rows are prefetched in register tiles
// of 8 frames.  The total thread-instruction count per frame is what the warp version spent on idle and duplicated lanes.
#define VIT_NT 256
__global__ void __launch_bounds__(VIT_NT, 2) k_pitch_viterbi(Clips c, PitchPass p) {
    __shared__ double s_delta[2][16], s_lf[2][16];
    const int tid = threadIdx.x, lane = tid & 31;
    const int clip = blockIdx.x;
    const int nF = p.nF[clip];
    if (nF < 1) return;
    const int f0 = p.fstart[clip];
    const PitchCfg& g = p.cfg[c.cls[clip]];
    const double corr = 0.01 / g.dt;
    const double jumpCost = g.jump_cost * corr, vuvCost = g.vuv_cost * corr;
    const int c2 = tid >> 4, c1 = tid & 15;                 // c2 == 15 and c1 == 15 are idle (MAXCAND = 15)
    const unsigned hmask = 0xffffu << (lane & 16);          // the half-warp that shares this c2

    constexpr int VT = 8;            // a tile has to cover a DRAM round trip: 4 frames did not (ncu: 27 % long-scoreboard stalls)
    double tsc[VT], tlf[VT], nsc[VT], nlf[VT];
    int tnc[VT], nnc[VT];
    auto load_tile = [&](int i0, double (&sc_)[VT], double (&lf_)[VT], int (&nc_)[VT]) {
#pragma unroll
        for (int u = 0; u < VT; u++) {
            const int i = i0 + u;
            sc_[u] = -1e300; lf_[u] = -1.0; nc_[u] = 0;
            if (i < nF) {
                const size_t fo = (size_t)(f0 + i);
                nc_[u] = p.ncand[fo];
                if (c2 < MAXCAND) { sc_[u] = p.cand_score[fo * MAXCAND + c2]; lf_[u] = p.cand_lf[fo * MAXCAND + c2]; }
            }
        }
    };
    load_tile(0, nsc, nlf, nnc);
    int cur = 0, ncPrev = 0;
    for (int i0 = 0; i0 < nF; i0 += VT) {
#pragma unroll
        for (int u = 0; u < VT; u++) { tsc[u] = nsc[u]; tlf[u] = nlf[u]; tnc[u] = nnc[u]; }
        if (i0 + VT < nF) load_tile(i0 + VT, nsc, nlf, nnc);
#pragma unroll
        for (int u = 0; u < VT; u++) {
            const int i = i0 + u;
            if (i >= nF) break;
            const size_t fo = (size_t)(f0 + i);
            const int nc = tnc[u];
            const double sc = tsc[u], clf = tlf[u];
            if (i == 0) {
                if (c1 == 0) { s_delta[cur][c2] = sc; s_lf[cur][c2] = clf; }
                ncPrev = nc;
                __syncthreads();
                continue;
            }
            const double pd = s_delta[cur][c1], plf = s_lf[cur][c1];
            const bool curVoiceless = clf < 0.0, prevVoiceless = plf < 0.0;
            double cost;
            if (curVoiceless) cost = prevVoiceless ? 0.0 : vuvCost;
            else cost = prevVoiceless ? vuvCost : jumpCost * fabs(plf - clf);
            const double value = pd - cost + sc;
            // candidates beyond ncPrev do not exist; the sequential scan starts from maximum = -1e30 and only takes larger values
            double best = (c1 < ncPrev && value > -1e30) ? value : -1e30;
            int place = c1;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(hmask, best, o);
                const int op = __shfl_xor_sync(hmask, place, o);
                if (ob > best || (ob == best && op < place)) { best = ob; place = op; }
            }
            if (best <= -1e30) place = 0;
            if (c1 == 0 && c2 < MAXCAND) {
                p.psi[fo * 16 + c2] = (uint8_t)place;
                s_delta[cur ^ 1][c2] = c2 < nc ? best : -1e300;
                s_lf[cur ^ 1][c2] = clf;
            }
            ncPrev = nc;
            cur ^= 1;
            __syncthreads();
        }
    }
    if (tid >= 32) return;
    // end of the most probable path: first maximum over the last frame's candidates
    int place = 0;
    {
        double best = s_delta[cur][0];
        for (int ci = 1; ci < ncPrev; ci++) {
            const double d = s_delta[cur][ci];
            if (d > best) { best = d; place = ci; }
        }
    }
    __syncwarp();
    // backtrack in tiles of 32 frames: lanes prefetch the psi rows, then the chain is resolved with shuffles
    for (int hi = nF - 1; hi >= 0; hi -= 32) {
        int lo = hi - 31 < 0 ? 0 : hi - 31;
        int myFrame = hi - lane;                                // lane 0 = frame hi
        unsigned long long r0 = 0, r1 = 0;
        if (myFrame >= lo && myFrame >= 1) {
            const unsigned long long* row = (const unsigned long long*)(p.psi + (size_t)(f0 + myFrame) * 16);
            r0 = row[0]; r1 = row[1];
        }
        int myPlace = 0;
        for (int l = 0; l <= hi - lo; l++) {
            if (lane == l) myPlace = place;
            // psi of frame (hi - l) at column `place`
            unsigned long long q0 = __shfl_sync(FULL_MASK, r0, l), q1 = __shfl_sync(FULL_MASK, r1, l);
            int fr = hi - l;
            if (fr >= 1) {
                unsigned long long qq = place < 8 ? q0 : q1;
                place = (int)((qq >> (8 * (place & 7))) & 0xff);
            }
        }
        if (myFrame >= lo) {
            size_t fo = (size_t)(f0 + myFrame);
            p.sel_f[fo] = p.cand_f[fo * MAXCAND + myPlace];
            p.sel_s[fo] = p.cand_s[fo * MAXCAND + myPlace];
        }
    }
}

void launch_pitch_viterbi(const Clips& c, const PitchPass& p, cudaStream_t s) {
    k_pitch_viterbi<<<c.n, VIT_NT, 0, s>>>(c, p);
}

// ------------------------------------------------------------------------------------------------ statistics
__device__ __forceinline__ bool voiced_f(double f, double ceiling) { return f > 0.0 && f < ceiling; }

// _pitch_values (mshds_extractor.py:143-162): z-score filtered mean of the wide pass decides the speaker class.
__global__ void __launch_bounds__(256) k_pitch_class(Clips c, PitchPass p) {
    __shared__ double red[32];
    const int clip = blockIdx.x;
    const int nF = p.nF[clip], f0 = p.fstart[clip];
    double s = 0.0, n = 0.0;
    for (int i = threadIdx.x; i < nF; i += blockDim.x) {
        double f = p.sel_f[f0 + i];
        if (f != 0.0) { s += f; n += 1.0; }
    }
    s = block_sum(s, red); n = block_sum(n, red);
    int cls = CLS_FALLBACK;
    if (n > 0.0) {
        double mean = s / n, v = 0.0;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) {
            double f = p.sel_f[f0 + i];
            if (f != 0.0) v += (f - mean) * (f - mean);
        }
        v = block_sum(v, red);
        double sd = sqrt(v / n);
        double s2 = 0.0, m = 0.0;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) {
            double f = p.sel_f[f0 + i];
            if (f != 0.0) {
                double z = (f - mean) / sd;
                if (fabs(z) <= 2.0) { s2 += f; m += 1.0; }
            }
        }
        s2 = block_sum(s2, red); m = block_sum(m, red);
        if (m > 0.0) cls = (s2 / m < 170.0) ? CLS_MALE : CLS_FEMALE;
    }
    if (threadIdx.x == 0) {
        c.cls[clip] = cls;
        if (cls == CLS_FALLBACK) atomicOr(&c.status[clip], ST_PITCHRANGE_FALLBACK);
    }
}
void launch_pitch_class(const Clips& c, const PitchPass& p, cudaStream_t s) { k_pitch_class<<<c.n, 256, 0, s>>>(c, p); }

// _extract_pitch (mshds_extractor.py:178-180): "Get mean 0 0 Hertz", "Get standard deviation 0 0 semitones"
__global__ void __launch_bounds__(256) k_pitch_stats(Clips c, PitchPass p) {
    __shared__ double red[32];
    const int clip = blockIdx.x;
    const int nF = p.nF[clip], f0 = p.fstart[clip];
    const double ceiling = p.cfg[c.cls[clip]].ceiling;
    double s = 0.0, st = 0.0, n = 0.0;
    for (int i = threadIdx.x; i < nF; i += blockDim.x) {
        double f = p.sel_f[f0 + i];
        if (voiced_f(f, ceiling)) { s += f; st += 12.0 * log2(f / 100.0); n += 1.0; }
    }
    s = block_sum(s, red); st = block_sum(st, red); n = block_sum(n, red);
    double mean_hz = DEVNAN, sd_st = DEVNAN;
    if (nF >= 1 && n > 0.0) mean_hz = s / n;
    if (nF >= 1 && n >= 2.0) {
        double mean_st = st / n, v = 0.0;
        for (int i = threadIdx.x; i < nF; i += blockDim.x) {
            double f = p.sel_f[f0 + i];
            if (voiced_f(f, ceiling)) { double d = 12.0 * log2(f / 100.0) - mean_st; v += d * d; }
        }
        v = block_sum(v, red);
        sd_st = sqrt(v / (n - 1.0));
    }
    if (threadIdx.x == 0) {
        c.feat[(size_t)clip * N_FEAT + 5] = mean_hz;
        c.feat[(size_t)clip * N_FEAT + 6] = sd_st;
        if (nF < 1) atomicOr(&c.status[clip], ST_PITCH);
    }
}
void launch_pitch_stats(const Clips& c, const PitchPass& p, cudaStream_t s) { k_pitch_stats<<<c.n, 256, 0, s>>>(c, p); }

// _extract_harmonicity (mshds_extractor.py:221-222): mean of 10 log10(r/(1-r)) over frames that are not -200 dB
__global__ void __launch_bounds__(256) k_hnr_mean(Clips c, PitchPass p) {
    __shared__ double red[32];
    const int clip = blockIdx.x;
    const int nF = p.nF[clip], f0 = p.fstart[clip];
    double s = 0.0, n = 0.0;
    for (int i = threadIdx.x; i < nF; i += blockDim.x) {
        double f = p.sel_f[f0 + i];
        if (f == 0.0) continue;
        double r = p.sel_s[f0 + i];
        double v = r <= 1e-15 ? -150.0 : r > 1.0 - 1e-15 ? 150.0 : 10.0 * log10(r / (1.0 - r));
        s += v; n += 1.0;
    }
    s = block_sum(s, red); n = block_sum(n, red);
    if (threadIdx.x == 0) {
        c.feat[(size_t)clip * N_FEAT + 9] = (nF >= 1 && n > 0.0) ? s / n : DEVNAN;
        if (nF < 1) atomicOr(&c.status[clip], ST_HNR);
    }
}
void launch_hnr_mean(const Clips& c, const PitchPass& p, cudaStream_t s) { k_hnr_mean<<<c.n, 256, 0, s>>>(c, p); }
