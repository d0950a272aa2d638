// k_acw.cu -- autocorrelation pitch frames (Sound_into_PitchFrame, AC_HANNING branch of fon/Sound_to_Pitch.cpp), the
// arithmetic behind to_pitch_ac at mshds_extractor.py:104 (M = 2048 stays on k_pitch_frames), :143, :178 / :355, :241, :270.
//
// B200 design (round 2; replaces the CTA-per-frame shared-memory FFT of k_pitch_frames<false> for M <= 1024):
//
//   * ONE WARP PER FRAME (half a warp for the 1024-point transforms of high-pitched speakers: two frames per warp).  The
//     packed real FFT, the power spectrum and the inverse FFT live in registers (fftreg.cuh): 32 complex points per lane,
//     one XOR-swizzled trip through shared memory per transform, the real-FFT untangle done with warp shuffles.  No
//     __syncthreads() inside a frame; the only block barriers are the two per staged span.
//   * SAMPLES ARRIVE BY TMA.  A CTA turn is 8 consecutive frames of one recording = one contiguous span of
//     W + 7 hops samples (int16: ~3 KB).  Thread 0 fetches the span of the NEXT turn with cp.async.bulk (1-D bulk tensor
//     copy, completion on an mbarrier) into the other half of a double buffer while the four warps work on the current
//     one; local mean, window multiply and local peak of all 8 frames read the staged span instead of going through L1
//     eight times.  Turns are handed out by an atomic counter (frames of low-pitched speakers cost twice as much).
//   * the candidate search (maxima, parabolic frequency, sinc-30 strength, Praat's slot rule) runs warp-synchronously on
//     the correlation the warp just produced (rs0 aliases the dead exchange buffer).
//
// Results are the same r[lag] = ac[lag] / (ac[0] windowR[lag]) as before up to the rounding of a different (exact) FFT
// operation order; everything downstream (refinement queue, candidates, path finder) is unchanged.
#include <cstdio>
#include <cstdlib>
#include "internal.h"
#include "common.cuh"
#include "fft.cuh"
#include "num.cuh"
#include "fftreg.cuh"
#include "fftwarp.cuh"
#include "stage.cuh"

#define ACW_WARPS 4
#define ACW_NT (32 * ACW_WARPS)
#define ACW_TURN 8
#define ACW_XCH_BYTES (1024 * 16)          // exchange buffer per warp: 1024 complex doubles
#define ACW_TW1024_BYTES (32 * 32 * 16)    // dynamic shared memory: [pass twiddles 1024][pass twiddles 512][exchange x warps][2 stages]
#define ACW_TW_BYTES (ACW_TW1024_BYTES + 32 * 16 * 16)

struct AcwParams : StageParams {
    const double2* twb512;            // [32][16] exp(-2 pi i j q / 512)
    const double2* twb1024;           // [32][32] exp(-2 pi i j q / 1024)
};
typedef StageSeg AcwSeg;

// thread 0: next segment of work -> *sg, and its bulk copy into `stage` (armed on `bar`)
__device__ __forceinline__ void acw_fetch(const Clips& c, const PitchPass& p, const AcwParams& A, int total, int nturn,
                                          int& cur_f, int& turn_end, AcwSeg* sg, unsigned char* stage, unsigned long long* bar) {
    if (cur_f >= turn_end) {
        const int turn = atomicAdd(p.turn_counter, 1);
        if (turn >= nturn) { sg->n = 0; sg->tma_bytes = 0; return; }
        cur_f = turn * ACW_TURN;
        turn_end = cur_f + ACW_TURN < total ? cur_f + ACW_TURN : total;
    }
    const int clip = find_segment(p.fstart, c.n, cur_f);
    const int clip_end = p.fstart[clip + 1];
    const int f1 = turn_end < clip_end ? turn_end : clip_end;
    const int cls = c.cls[clip];
    const PitchCfg& g = p.cfg[cls];
    sg->f0 = cur_f; sg->n = f1 - cur_f; sg->clip = clip; sg->cls = cls; sg->k0 = cur_f - p.fstart[clip];
    const long long base = c.off[clip], nx = c.off[clip + 1] - base;
    // samples the frames k0 .. k0 + n - 1 touch: window [right - half, left + half] and local-mean range [right - period, left + period]
    const double x1 = c.x1[clip];
    const double tA = p.t1[clip] + (double)sg->k0 * g.dt, tB = p.t1[clip] + (double)(sg->k0 + sg->n - 1) * g.dt;
    const long long leftA = x_to_low(x1, c.dx, tA), leftB = x_to_low(x1, c.dx, tB);
    const int reach = g.halfnsamp_window > g.nsamp_period ? g.halfnsamp_window : g.nsamp_period;
    long long sA = leftA + 1 - reach, sB = leftB + reach;
    if (sA < 1) sA = 1;
    if (sB > nx) sB = nx;
    if (sB < sA) sB = sA - 1;
    stage_issue(A, sg, base, sA, (int)(sB - sA + 1), stage, bar);
    cur_f = f1;
}

// ------------------------------------------------------------------------------------------------ candidates (warp)
struct WCand {
    double* rs0;        // [2Bs+1] symmetric correlation, rs0[Bs + i] = r[i]
    double* pk_f; double* pk_s; double* pk_key; int* pk_lag;
    unsigned* masks;
    double* cf; double* cs; double* ckey; int* cimax;           // [16] 1-based slots, first analysis
    double* cf2; double* cs2; double* ckey2; int* cimax2;       // second analysis (dual voicing threshold)
    int* s_int;
    int pkcap;
};

__device__ __forceinline__ int w_insert_candidates(const PitchCfg& g, const WCand& S, const double* r, int nmax, double thr,
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

// Sound_into_PitchFrame first pass for one frame, executed by one warp (same arithmetic as find_candidates in k_pitch.cu).
// Returns ncand | ncand2 << 8.
__device__ __forceinline__ int w_find_candidates(const PitchCfg& g, double dx, const WCand& S, int B, int Bs,
                                                 const double2* __restrict__ tw, double vt2) {
    const int lane = threadIdx.x & 31;
    const double thr1 = 0.5 * g.vt;
    const double thr = (vt2 >= 0.0 && 0.5 * vt2 < thr1) ? 0.5 * vt2 : thr1;
    const int upper = g.maximumLag < B ? g.maximumLag : B;      // i < maximumLag && i < brent_ixmax
    int nlag = upper - 2;
    if (nlag < 0) nlag = 0;
    const int nrounds = (nlag + 31) / 32;
    const double* r = S.rs0 + Bs;
    int n = 0;
    for (int round = 0; round < nrounds; round++) {
        const int i = 2 + round * 32 + lane;
        bool flag = false;
        if (i < 2 + nlag) {
            const double ri = r[i];
            flag = ri > thr && ri > r[i - 1] && ri >= r[i + 1];
        }
        const unsigned m = __ballot_sync(FULL_MASK, flag);
        // ordered compaction: every lane knows the running count
        if (flag) {
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            if (pos < S.pkcap) S.pk_lag[pos] = i;
        }
        n += __popc(m);
    }
    if (n > S.pkcap) n = S.pkcap;
    __syncwarp();
    const int nmax = n;
    const double* y1 = S.rs0 - 1;
    const int ny = 2 * Bs + 1;
    for (int m = 0; m < nmax; m++) {
        const int i = S.pk_lag[m];
        const double dr = 0.5 * (r[i + 1] - r[i - 1]), d2r = 2 * r[i] - r[i - 1] - r[i + 1];
        const double freq = 1.0 / dx / (i + dr / d2r);
        const double x = 1.0 / dx / freq + (double)(Bs + 1);
        double strength = sinc_interp_warp(y1, ny, x, 30, lane, tw);
        if (strength > 1.0) strength = 1.0 / strength;
        if (lane == 0) {
            S.pk_f[m] = freq;
            S.pk_s[m] = strength;
            S.pk_key[m] = strength - g.octave_cost * log2(g.floor_hz / freq);
        }
    }
    __syncwarp();
    if (lane == 0) S.s_int[1] = w_insert_candidates(g, S, r, nmax, thr1, S.cf, S.cs, S.ckey, S.cimax);
    if (lane == 1 && vt2 >= 0.0) S.s_int[2] = w_insert_candidates(g, S, r, nmax, 0.5 * vt2, S.cf2, S.cs2, S.ckey2, S.cimax2);
    __syncwarp();
    return S.s_int[1] | ((vt2 >= 0.0 ? S.s_int[2] : 0) << 8);
}

// ------------------------------------------------------------------------------------------------ the frame, per lane
// The correlation of one frame.  L (16 or 32) lanes per frame is a RUN-TIME value: both transform sizes share one copy of the
// code (fr_slot<16> == fr_slot<32> == 5-bit reversal; pass B of the 512-point case is the 32-point butterfly network minus its
// first stage = two 16-point transforms), and forward and inverse transform share it too (inverse = conj FFT conj, the
// conjugations folded into the untangle step and the read-out).  What is left is ~2,500 instructions of straight-line
// code executed twice per frame instead of four different unrolled transforms per size: the first version of this kernel
// spent 35 % of its stall samples waiting for instruction fetch (ncu: stalled_no_instructions).
__device__ __noinline__ void acw_frame(const Clips& c, const PitchPass& p, const AcwParams& A, const AcwSeg& sg,
                                       const unsigned char* st, unsigned char* xw /* warp exchange region */, int fi,
                                       bool active, int L, const double2* __restrict__ tw) {
    const int lane = threadIdx.x & 31, j = lane & (L - 1), gidx = L == 32 ? 0 : lane >> 4;
    const unsigned gmask = L == 32 ? FULL_MASK : (0xffffu << (16 * gidx));
    const PitchCfg& g = p.cfg[sg.cls];
    const double dx = c.dx;
    const int W = g.nsamp_window, B = g.brent_ixmax;
    const int esz = A.esz;
    double2* xch = (double2*)(xw + (size_t)gidx * (ACW_XCH_BYTES / 2));     // this frame's exchange region (L = 16: half)
    extern __shared__ __align__(16) unsigned char acw_smem[];               // the kernel's dynamic shared memory: tables first
    const FwTwShared twf((const double2*)(acw_smem + (L == 32 ? 0 : ACW_TW1024_BYTES)), j, L);

    const double x1 = c.x1[sg.clip];
    const double t = p.t1[sg.clip] + (double)(sg.k0 + fi) * g.dt;
    const long long leftSample = x_to_low(x1, dx, t), rightSample = leftSample + 1;
    const int eoff = (int)(-sg.sA) + sg.shift;            // staged element (incl. alignment shift) of 1-based clip sample s: s + eoff

    // ---- local mean over one longest period to both sides
    double acc = 0.0;
    if (active) {
        const long long s0 = rightSample - g.nsamp_period, s1 = leftSample + g.nsamp_period;
        for (long long i = s0 + j; i <= s1; i += L) acc += staged(st, esz, (int)i + eoff);
    }
    const double localMean = group_sum(acc, gmask, L) / (double)(2 * g.nsamp_period);

    // ---- frame (s - mean) * Hanning window into registers, packed z[n] = x[2n] + i x[2n+1], n = j + L k; local peak
    double2 a[32];
    double lp = 0.0;
    {
        const long long startSample = rightSample - g.halfnsamp_window;
        int pk0 = g.halfnsamp_window + 1 - g.halfnsamp_period; if (pk0 < 1) pk0 = 1;
        int pk1 = g.halfnsamp_window + g.halfnsamp_period; if (pk1 > W) pk1 = W;
        const int ebase = (int)startSample + eoff;                        // staged element of frame sample m = 0
        fr_static_for<0, 32>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            const int m = 2 * (j + L * k);
            double v0 = 0.0, v1 = 0.0;
            if (active && m < W) {                                        // W is even: m + 1 < W as well
                const double2 w = __ldg((const double2*)(g.window + m));
                v0 = (staged(st, esz, ebase + m) - localMean) * w.x;
                v1 = (staged(st, esz, ebase + m + 1) - localMean) * w.y;
                if (m + 1 >= pk0 && m + 1 <= pk1) lp = fmax(lp, fabs(v0));
                if (m + 2 >= pk0 && m + 2 <= pk1) lp = fmax(lp, fabs(v1));
            }
            a[k] = make_double2(v0, v1);
        });
    }
    for (int o = L >> 1; o > 0; o >>= 1) lp = fmax(lp, __shfl_xor_sync(gmask, lp, o));
    const double localPeak = lp;
    const double globalPeak = active ? c.gpeak[sg.clip] : 1.0;
    const double intensity = localPeak > globalPeak ? 1.0 : localPeak / globalPeak;

    // ---- autocorrelation = inverse FFT of |FFT|^2: the SAME transform code runs twice
    const double2 wj = __ldg(tw + j * (TW_N / (64 * L)));                 // exp(-2 pi i j / N), N = 64 L
    fw_roundtrip(a, xch, lane, j, L, twf, wj, FwIdentity());
    // a[brev5(r)] = conj(y[n]), n = j + L r, y[n] = ac[2n] + i ac[2n+1].  Lags 0..B go through shared memory in natural
    // order; the normalised correlation r[i] = ac[i] / (ac[0] windowR[i]) is written by a compact loop (coalesced rows)
    double* acs = (double*)xch;
    fr_static_for<0, 12>([&](auto rc) {                                    // B / 2 < 12 L for every configuration (checked by the launcher)
        constexpr int r = decltype(rc)::value;
        const int n = j + L * r;
        if (2 * n <= B) {
            const double2 v = a[fr_brev(r, 5)];
            *(double2*)(acs + 2 * n) = make_double2(v.x, -v.y);
        }
    });
    __syncwarp();
    if (active) {
        const int f = sg.f0 + fi;
        double* rrow = p.rbuf + (size_t)f * p.rstride;
        const double ac0 = acs[0];
        // four window-autocorrelation values in flight per trip (two warps per scheduler cannot hide one L1/L2 round trip per lag)
        for (int i0 = j; i0 <= B; i0 += 4 * L) {
            double wr[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int i = i0 + u * L; wr[u] = i <= B ? __ldg(g.windowR + i) : 1.0; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * L;
                if (i <= B) rrow[i] = i == 0 ? 1.0 : acs[i] / (ac0 * wr[u]);
            }
        }
        if (j == 0) {
            p.inten[f] = intensity;
            p.ncand[f] = localPeak != 0.0 ? 1 : 0;        // hand-over to k_ac_candidates: "the frame has a local peak"
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------ kernel
template <int OCC>
__global__ void __launch_bounds__(ACW_NT, OCC) k_ac_frames_w(const __grid_constant__ Clips c, const __grid_constant__ PitchPass p,
                                                             const __grid_constant__ AcwParams A, const double2* __restrict__ tw) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char* xch_all = smem + ACW_TW_BYTES;                         // ACW_WARPS x 16 KB
    unsigned char* stage0 = xch_all + ACW_WARPS * ACW_XCH_BYTES;          // 2 x stage_bytes
    __shared__ __align__(8) unsigned long long bars[2];
    __shared__ AcwSeg segs[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = p.fstart[c.n];
    const int nturn = (total + ACW_TURN - 1) / ACW_TURN;
    int cur_f = 0, turn_end = 0;                                          // thread 0's position in its current turn
    if (tid == 0) mbar_init_pair(bars);
    fw_stage_twiddles<ACW_NT>((double2*)smem, A.twb1024, 32);            // visible after the first barrier of the loop below
    fw_stage_twiddles<ACW_NT>((double2*)(smem + ACW_TW1024_BYTES), A.twb512, 16);
    __syncthreads();
    if (tid == 0) acw_fetch(c, p, A, total, nturn, cur_f, turn_end, &segs[0], stage0, &bars[0]);
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    for (;;) {
        __syncthreads();                      // segs[buf] is published; every warp is done with the other stage buffer
        const AcwSeg& sg = segs[buf];            // stays valid for this iteration: thread 0 only writes segs[buf ^ 1]
        if (sg.n == 0) break;
        unsigned char* st = stage0 + (size_t)buf * A.stage_bytes;
        if (tid == 0)                         // prefetch the next segment while this one is processed
            acw_fetch(c, p, A, total, nturn, cur_f, turn_end, &segs[buf ^ 1], stage0 + (size_t)(buf ^ 1) * A.stage_bytes, &bars[buf ^ 1]);
        stage_complete<ACW_NT>(A, sg, st, &bars[buf], buf == 0 ? phase0 : phase1);
        unsigned char* xw = xch_all + (size_t)warp * ACW_XCH_BYTES;
        const int L = p.cfg[sg.cls].M == 1024 ? 32 : 16;
        const int per_warp = 32 / L;                                     // frames a warp handles at once
        for (int fr0 = 0; fr0 < sg.n; fr0 += per_warp * ACW_WARPS) {
            const int first = fr0 + per_warp * warp;
            if (first >= sg.n) break;                                    // warp-uniform
            const int fi = first + (L == 32 ? 0 : lane >> 4);
            acw_frame(c, p, A, sg, st, xw, fi, fi < sg.n, L, tw);
        }
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------ candidate kernel
// Sound_into_PitchFrame, first pass, for the frames whose correlation rows k_ac_frames_w just wrote: one WARP per frame at
// high occupancy (this part is a chain of dependent shuffles and a serial slot insertion -- latency, not arithmetic -- and
// used to sit in the register-heavy transform kernel, where only 8 warps per SM could hide it).
#define ACC_WARPS 4
__global__ void __launch_bounds__(32 * ACC_WARPS) k_ac_candidates(const __grid_constant__ Clips c, const __grid_constant__ PitchPass p,
                                                                  const double2* __restrict__ tw, int region_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* xf = smem + (size_t)warp * region_bytes;
    const int total = p.fstart[c.n];
    const double dx = c.dx;
    const bool dual = p.dual_cand_f != nullptr;
    for (int f = blockIdx.x * ACC_WARPS + warp; f < total; f += gridDim.x * ACC_WARPS) {
        const int clip = find_segment(p.fstart, c.n, f);
        const PitchCfg& g = p.cfg[c.cls[clip]];
        const int B = g.brent_ixmax;
        const int Bs = B < g.maximumLag + 32 ? B : g.maximumLag + 32;
        WCand S;
        {
            unsigned char* q = xf;
            S.rs0 = (double*)q; q += ((size_t)(2 * Bs + 2) * 8 + 15) & ~(size_t)15;
            int pkcap = g.maximumLag / 2 + 2;
            if (pkcap > 320) pkcap = 320;
            S.pkcap = pkcap;
            S.pk_f = (double*)q; q += (size_t)pkcap * 8;
            S.pk_s = (double*)q; q += (size_t)pkcap * 8;
            S.pk_key = (double*)q; q += (size_t)pkcap * 8;
            S.cf = (double*)q; q += 16 * 8; S.cs = (double*)q; q += 16 * 8; S.ckey = (double*)q; q += 16 * 8;
            S.cf2 = (double*)q; q += 16 * 8; S.cs2 = (double*)q; q += 16 * 8; S.ckey2 = (double*)q; q += 16 * 8;
            S.pk_lag = (int*)q; q += ((size_t)pkcap * 4 + 15) & ~(size_t)15;
            S.cimax = (int*)q; q += 16 * 4; S.cimax2 = (int*)q; q += 16 * 4;
            S.s_int = (int*)q; q += 16;
            S.masks = (unsigned*)q;
        }
        const double inten = p.inten[f];
        const bool has_peak = p.ncand[f] != 0;
        const double* rrow = p.rbuf + (size_t)f * p.rstride;
        __syncwarp();
        for (int i = lane; i <= Bs; i += 32) { const double v = rrow[i]; S.rs0[Bs + i] = v; S.rs0[Bs - i] = v; }
        __syncwarp();
        int ncand = 1, ncand2 = 1;
        if (has_peak) {
            const int nn = w_find_candidates(g, dx, S, B, Bs, tw, dual ? p.dual_vt : -1.0);
            ncand = nn & 0xff;
            if (dual) ncand2 = nn >> 8;
        } else {
            if (lane == 0) { S.cf[1] = 0.0; S.cs[1] = 0.0; S.cimax[1] = 0; S.cf2[1] = 0.0; S.cs2[1] = 0.0; S.cimax2[1] = 0; }
            __syncwarp();
        }
        // candidates leave the SM (lanes 0..14: the analysis itself, lanes 16..30: the dual one); see k_pitch_frames for the
        // exactness argument of the two skip rules
        const int set = lane >> 4, ctid = lane & 15;
        if (ctid < MAXCAND && (set == 0 || dual)) {
            const double vt = set == 0 ? g.vt : p.dual_vt;
            double uvs = g.sil <= 0 ? 0.0 : 2.0 - inten / (g.sil / (1.0 + vt));
            uvs = vt + (uvs > 0 ? uvs : 0);
            const bool frame_stays_unvoiced = uvs > 1.0 + 2.0 * g.vuv_cost * (0.01 / g.dt) + 1e-9;
            const int nc = set == 0 ? ncand : ncand2;
            const double* scf = set == 0 ? S.cf : S.cf2;
            const double* scs = set == 0 ? S.cs : S.cs2;
            const int* sci = set == 0 ? S.cimax : S.cimax2;
            const int ci = ctid + 1;
            double fr = 0.0, stn = 0.0;
            int im = 0;
            if (ci <= nc) { fr = scf[ci]; stn = scs[ci]; im = sci[ci]; }
            const size_t o2 = (size_t)f * MAXCAND + ctid;
            if (set == 0) { p.cand_f[o2] = fr; p.cand_s[o2] = stn; p.cand_imax[o2] = (unsigned short)im; }
            else { p.dual_cand_f[o2] = fr; p.dual_cand_s[o2] = stn; p.dual_cand_imax[o2] = (unsigned short)im; }
            const bool live = ci >= 2 && ci <= nc && (1.0 / dx / (double)(im + 1) < g.ceiling) && !frame_stays_unvoiced;
            if (live) {
                if (set == 0) { int slot = atomicAdd(p.qcount, 1); p.queue[slot] = f * 16 + ctid; }
                else { int slot = atomicAdd(p.dual_qcount, 1); p.dual_queue[slot] = f * 16 + ctid; }
            }
        }
        if (lane == 0) {
            p.ncand[f] = (uint8_t)ncand;
            if (dual) { p.dual_ncand[f] = (uint8_t)ncand2; p.dual_inten[f] = inten; }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ launcher
// scratch of one warp of k_ac_candidates, the largest over the speaker classes
static size_t acc_region_bytes(const PitchPass& p) {
    size_t region = 0;
    for (int k = 0; k < 3; k++) {
        const PitchCfg& g = p.cfg[k];
        const int Bs = g.brent_ixmax < g.maximumLag + 32 ? g.brent_ixmax : g.maximumLag + 32;
        int pkcap = g.maximumLag / 2 + 2;
        if (pkcap > 320) pkcap = 320;
        const size_t need = (((size_t)(2 * Bs + 2) * 8 + 15) & ~(size_t)15) + (size_t)pkcap * 24 + 6 * 128 + (((size_t)pkcap * 4 + 15) & ~(size_t)15) +
                            2 * 64 + 16 + 64;
        if (need > region) region = need;
    }
    return (region + 127) & ~(size_t)127;
}

// candidates: one warp per frame, as many warps per SM as the scratch allows (also serves the cross-correlation frames of
// k_ccs.cu); returns false when the scratch does not fit
bool launch_ac_candidates(const Clips& c, const PitchPass& p, const double2* tw, int max_frames_hint, cudaStream_t s) {
    const size_t region = acc_region_bytes(p);
    const size_t smem2 = region * ACC_WARPS;
    if (smem2 > 200 * 1024) return false;
    cudaFuncSetAttribute(k_ac_candidates, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    cudaFuncSetAttribute(k_ac_candidates, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    int occ2 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_ac_candidates, 32 * ACC_WARPS, smem2);
    if (occ2 < 1) occ2 = 1;
    int grid2 = sm_count() * occ2;
    const int need_blocks = (max_frames_hint + ACC_WARPS - 1) / ACC_WARPS;
    if (max_frames_hint > 0 && grid2 > need_blocks) grid2 = need_blocks;
    if (grid2 < 1) grid2 = 1;
    k_ac_candidates<<<grid2, 32 * ACC_WARPS, smem2, s>>>(c, p, tw, (int)region);
    return true;
}

// returns false when this pass cannot run on the warp kernel (transform larger than 1024 points, scratch does not fit)
bool launch_ac_frames_warp(const Clips& c, const PitchPass& p, const double2* tw, const double2* twb512, const double2* twb1024,
                           long long total_elems, int max_frames_hint, cudaStream_t s) {
    int span = 0;
    for (int k = 0; k < 3; k++) {
        const PitchCfg& g = p.cfg[k];
        if (g.method != 0 || (g.M != 512 && g.M != 1024)) return false;
        const int Lk = g.M / 32;
        if (g.brent_ixmax / 2 >= 12 * Lk || (g.nsamp_window & 1)) return false;           // read-out loop covers r < 12
        const int reach = g.halfnsamp_window > g.nsamp_period ? g.halfnsamp_window : g.nsamp_period;
        const int hop = (int)ceil(g.dt / c.dx) + 1;
        const int sp = 2 * reach + 2 + (ACW_TURN - 1) * hop + 16;
        if (sp > span) span = sp;
    }
    const size_t region = acc_region_bytes(p);
    AcwParams A;
    A.esz = c.pcm.p64 ? 8 : 2;
    A.pcm_bytes = c.pcm.p64 ? (const unsigned char*)c.pcm.p64 : (const unsigned char*)c.pcm.p16;
    A.total_elems = total_elems;
    A.stage_bytes = (span * A.esz + 32 + 127) & ~127;
    A.twb512 = twb512; A.twb1024 = twb1024;
    const size_t smem = (size_t)ACW_TW_BYTES + (size_t)ACW_WARPS * ACW_XCH_BYTES + 2 * (size_t)A.stage_bytes;
    if (smem > 110 * 1024 || region * ACC_WARPS > 200 * 1024) return false;
    cudaMemsetAsync(p.qcount, 0, sizeof(int), s);
    cudaMemsetAsync(p.turn_counter, 0, sizeof(int), s);
    if (p.dual_cand_f) cudaMemsetAsync(p.dual_qcount, 0, sizeof(int), s);
    static int occ_want = 0;
    if (!occ_want) { const char* e = getenv("MSHDS_ACW_OCC"); occ_want = e && atoi(e) == 3 ? 3 : 2; }     // development switch (A/B)
    const int nturn = (max_frames_hint + ACW_TURN - 1) / ACW_TURN;
#define ACW_LAUNCH(O) \
    do { \
        cudaFuncSetAttribute(k_ac_frames_w<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        cudaFuncSetAttribute(k_ac_frames_w<O>, cudaFuncAttributePreferredSharedMemoryCarveout, 100); \
        int occ = 0; \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ac_frames_w<O>, ACW_NT, smem); \
        if (occ < 1) occ = 1; \
        int grid = sm_count() * occ; \
        if (max_frames_hint > 0 && grid > nturn) grid = nturn; \
        if (grid < 1) grid = 1; \
        k_ac_frames_w<O><<<grid, ACW_NT, smem, s>>>(c, p, A, tw); \
    } while (0)
    if (occ_want == 3) ACW_LAUNCH(3); else ACW_LAUNCH(2);
#undef ACW_LAUNCH
    launch_ac_candidates(c, p, tw, max_frames_hint, s);
    return true;
}
