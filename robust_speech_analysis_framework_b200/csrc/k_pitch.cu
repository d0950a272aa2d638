// k_pitch.cu -- Boersma (1993) pitch analysis on the GPU: autocorrelation (FFT) and forward cross-correlation frames,
// candidate picking with sinc refinement, Viterbi path finder, and the per-clip pitch statistics.
//
// Replaces the Praat calls behind mshds_extractor.py:104 (speech-rate pitch), :143 (_pitch_values), :178/:355
// (to_pitch_ac), :221 (to_harmonicity_cc), :241 (pitch inside "To Ltas (pitch-corrected)"), :270 (CPP pitch),
// :320 (to_pitch_cc) -- i.e. fon/Sound_to_Pitch.cpp Sound_to_Pitch_any + Sound_into_PitchFrame and
// fon/Pitch.cpp Pitch_pathFinder.
//
// One CTA per frame (persistent grid-stride loop over the flattened frame list of all clips): the frame is staged in
// shared memory, transformed with the packed real FFT of fft.cuh (AC) or correlated directly (FCC), candidates are
// refined warp-per-candidate, and only the <=15 candidates leave the SM.
#include "internal.h"
#include "common.cuh"
#include "fft.cuh"
#include "num.cuh"

#define NTHR 256
#define NWARP (NTHR / 32)
#define MAXPK 320

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
    if (ok) ok = short_term_analysis(nx, c.dx, 0.5 * c.dx, g.grid_window, g.dt, &nf, &t1) != 0;
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
    double* rs0;        // [2B+1] symmetric correlation, rs0[B+i] = r[i]
    double* pk_f;       // [MAXPK]
    double* pk_s;
    double* pk_key;
    int* pk_lag;
    unsigned* masks;    // [ceil(maxlag/32)+8]
    double* cf;         // [maxn+1] 1-based candidate slots
    double* cs;
    double* ckey;
    int* cimax;
    int* s_int;         // [4]: n maxima, ncand
};

// Sound_into_PitchFrame, second half: local maxima of r -> candidate slots -> sinc refinement.  Returns ncand (>=1).
__device__ __forceinline__ int find_and_refine(const PitchCfg& g, double dx, double ceiling, const CandScratch& S, int B,
                                               bool refine_all) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double thr = 0.5 * g.vt;
    int upper = g.maximumLag < B ? g.maximumLag : B;      // i < maximumLag && i < brent_ixmax
    int nlag = upper - 2;                                   // lags 2 .. upper-1
    if (nlag < 0) nlag = 0;
    int nrounds = (nlag + NTHR - 1) / NTHR;
    const double* r = S.rs0 + B;                            // r[i], i in [-B, B]
    for (int round = 0; round < nrounds; round++) {
        int i = 2 + round * NTHR + tid;
        bool flag = false;
        if (i < 2 + nlag) {
            double ri = r[i];
            flag = ri > thr && ri > r[i - 1] && ri >= r[i + 1];
        }
        unsigned m = __ballot_sync(FULL_MASK, flag);
        if (lane == 0) S.masks[round * NWARP + warp] = m;
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int w = 0; w < nrounds * NWARP; w++) {
            unsigned m = S.masks[w];
            while (m) {
                int b = __ffs(m) - 1;
                m &= m - 1;
                if (n < MAXPK) S.pk_lag[n++] = 2 + w * 32 + b;
            }
        }
        S.s_int[0] = n;
    }
    __syncthreads();
    const int nmax = S.s_int[0];
    const double* y1 = S.rs0 - 1;                           // 1-based view: y1[j] = r[j - B - 1]
    const int ny = 2 * B + 1;
    // first pass: parabolic frequency, sinc(30) strength
    for (int m = warp; m < nmax; m += NWARP) {
        int i = S.pk_lag[m];
        double dr = 0.5 * (r[i + 1] - r[i - 1]), d2r = 2 * r[i] - r[i - 1] - r[i + 1];
        double freq = 1.0 / dx / (i + dr / d2r);
        double x = 1.0 / dx / freq + (double)(B + 1);
        double strength = sinc_interp_warp(y1, ny, x, 30, lane);
        if (strength > 1.0) strength = 1.0 / strength;
        if (lane == 0) {
            S.pk_f[m] = freq;
            S.pk_s[m] = strength;
            S.pk_key[m] = strength - g.octave_cost * log2(g.floor_hz / freq);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int ncand = 1;
        S.cf[1] = 0.0; S.cs[1] = 0.0; S.cimax[1] = 0;
        for (int m = 0; m < nmax; m++) {
            int place = 0;
            if (ncand < g.maxn) {
                place = ++ncand;
            } else {
                double weakest = 2;
                for (int iweak = 2; iweak <= g.maxn; iweak++) {
                    double ls = S.ckey[iweak];
                    if (ls < weakest) { weakest = ls; place = iweak; }
                }
                if (S.pk_key[m] <= weakest) place = 0;
            }
            if (place) {
                S.cf[place] = S.pk_f[m]; S.cs[place] = S.pk_s[m]; S.ckey[place] = S.pk_key[m]; S.cimax[place] = S.pk_lag[m];
            }
        }
        S.s_int[1] = ncand;
    }
    __syncthreads();
    const int ncand = S.s_int[1];
    // second pass: maximise the sinc(70/700) interpolation with Brent.  A candidate whose refined lag cannot fall
    // below fs/ceiling (lag <= imax+1) is voiceless for the path finder whatever its refined values: skip it.
    for (int ci = 2 + warp; ci <= ncand; ci += NWARP) {
        int imax = S.cimax[ci];
        bool live = refine_all || (1.0 / dx / (double)(imax + 1) < ceiling);
        if (!live) continue;
        double xmid;
        double ymid = improve_extremum_warp(y1, ny, imax + B + 1, S.cf[ci] > 0.3 / dx ? PEAK_SINC700 : PEAK_SINC70, &xmid,
                                            true, lane);
        xmid -= (double)(B + 1);
        if (ymid > 1.0) ymid = 1.0 / ymid;
        if (lane == 0) { S.cf[ci] = 1.0 / dx / xmid; S.cs[ci] = ymid; }
    }
    __syncthreads();
    return ncand;
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
    int a, rs, pkf, pks, pkkey, cf, cs, ckey, red, part, pklag, cimax, masks, sint, fi, total;
    int part_stride;    // CC: doubles per partial-sum row (>= maximumLag)
};

template <bool IS_CC>
__global__ void __launch_bounds__(NTHR) k_pitch_frames(Clips c, PitchPass p, const double2* __restrict__ tw, FrameSmem L) {
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
    double* red = (double*)(smem + L.red);
    double* part = (double*)(smem + L.part);
    S.pk_lag = (int*)(smem + L.pklag);
    S.cimax = (int*)(smem + L.cimax);
    S.masks = (unsigned*)(smem + L.masks);
    S.s_int = (int*)(smem + L.sint);
    FrameInfo* fi = (FrameInfo*)(smem + L.fi);
    const int PS = L.part_stride;

    const int tid = threadIdx.x;
    const int total = p.fstart[c.n];
    const double dx = c.dx;
    const double x1 = 0.5 * dx;

    for (int f = blockIdx.x; f < total; f += gridDim.x) {
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
        const PitchCfg& g = p.cfg[fi->cls];
        const int16_t* pcm = c.pcm + fi->base;            // pcm[i-1] is 1-based sample i
        const long long nx = fi->nx;
        const double t = p.t1[clip] + (double)fi->k * g.dt;
        const long long leftSample = x_to_low(x1, dx, t), rightSample = leftSample + 1;
        const double globalPeak = c.gpeak[clip];
        const int B = g.brent_ixmax;
        const int W = g.nsamp_window;

        // local mean over one longest period to both sides
        double acc = 0.0;
        {
            long long s0 = rightSample - g.nsamp_period, s1 = leftSample + g.nsamp_period;
            for (long long i = s0 + tid; i <= s1; i += NTHR) acc += samp(pcm, i - 1);
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
            for (int m = tid; m < Nfft; m += NTHR) {
                double v = 0.0;
                if (m < W) {
                    v = (samp(pcm, startSample + m - 1) - localMean) * __ldg(g.window + m);
                    int j = m + 1;
                    if (j >= pk0 && j <= pk1) lp = fmax(lp, fabs(v));
                }
                ar[m] = v;
            }
        } else {
            for (int m = tid; m < W; m += NTHR) {
                int j = m + 1;
                if (j >= pk0 && j <= pk1) lp = fmax(lp, fabs(samp(pcm, startSample + m - 1) - localMean));
            }
        }
        const double localPeak = block_max(lp, red);
        const double intensity = localPeak > globalPeak ? 1.0 : localPeak / globalPeak;
        __syncthreads();

        if (!IS_CC) {
            fft_dif<-1>(a, g.M, tw);
            packed_power_to_inverse_input(a, g.M, g.logM, tw, IdentityF(), (double*)nullptr);
            fft_dit<+1>(a, g.M, tw);
            const double* ac = (const double*)a;            // ac[i] natural order
            const double ac0 = ac[0];
            for (int i = tid; i <= B; i += NTHR) {
                double v = i == 0 ? 1.0 : ac[i] / (ac0 * __ldg(g.windowR + i));
                S.rs0[B + i] = v;
                S.rs0[B - i] = v;
            }
        } else {
            // forward cross-correlation
            const double startTime = t - 0.5 * (1.0 / g.floor_hz + g.dt_window);
            long long startS = x_to_low(x1, dx, startTime);
            if (startS < 1) startS = 1;
            long long localSpan = g.maximumLag + W;
            if (localSpan > nx + 1 - startS) localSpan = nx + 1 - startS;
            const int localMaximumLag = (int)(localSpan - W);
            // xs[j-1] = s[startS-1+j] - localMean, j = 1..localSpan
            for (int j = tid; j < (int)localSpan; j += NTHR) xs[j] = samp(pcm, startS - 1 + j) - localMean;
            for (int i = tid; i < 2 * B + 1; i += NTHR) S.rs0[i] = 0.0;
            __syncthreads();
            // sumx2 over the first window
            double sx = 0.0;
            for (int j = tid; j < W; j += NTHR) sx = fma(xs[j], xs[j], sx);
            const double sumx2 = block_sum(sx, red);
            // products: work item = (lag, quarter of the window)
            const int L = localMaximumLag > 0 ? localMaximumLag : 0;
            const int q = (W + 3) / 4;
            for (int wi = tid; wi < 4 * L; wi += NTHR) {
                int lag = wi % L + 1, ch = wi / L;
                int j0 = ch * q, j1 = j0 + q < W ? j0 + q : W;
                double pr = 0.0, sy = 0.0;
                const double* xa = xs + j0;
                const double* xb = xs + j0 + lag;
                for (int j = 0; j < j1 - j0; j++) {
                    double yb = xb[j];
                    pr = fma(xa[j], yb, pr);
                    sy = fma(yb, yb, sy);
                }
                part[ch * PS + (lag - 1)] = pr;
                part[(4 + ch) * PS + (lag - 1)] = sy;
            }
            __syncthreads();
            for (int lag = 1 + tid; lag <= L; lag += NTHR) {
                double pr = (part[lag - 1] + part[PS + lag - 1]) + (part[2 * PS + lag - 1] + part[3 * PS + lag - 1]);
                double sy = (part[4 * PS + lag - 1] + part[5 * PS + lag - 1]) + (part[6 * PS + lag - 1] + part[7 * PS + lag - 1]);
                double v = pr / sqrt(sumx2 * sy);
                S.rs0[B + lag] = v;
                S.rs0[B - lag] = v;
            }
            if (tid == 0) S.rs0[B] = 1.0;
        }
        __syncthreads();

        int ncand = 1;
        if (localPeak != 0.0) {
            ncand = find_and_refine(g, dx, g.ceiling, S, B, p.hnr_mode != 0);
        } else if (tid == 0) {
            S.cf[1] = 0.0; S.cs[1] = 0.0;
        }
        __syncthreads();

        // Pitch_pathFinder local scores
        double unvoicedStrength = g.sil <= 0 ? 0.0 : 2.0 - intensity / (g.sil / (1.0 + g.vt));
        unvoicedStrength = g.vt + (unvoicedStrength > 0 ? unvoicedStrength : 0);
        if (p.hnr_mode) {
            // all path costs are zero: the path is the per-frame first maximum of the local scores
            if (tid == 0) {
                double best = unvoicedStrength, bestS = DEVNAN, bestF = 0.0;
                for (int ci = 2; ci <= ncand; ci++) {
                    double fr = S.cf[ci];
                    bool voiceless = !(fr > 0.0 && fr < g.ceiling);
                    double sc = voiceless ? unvoicedStrength : S.cs[ci];
                    if (sc > best) { best = sc; bestS = voiceless ? DEVNAN : S.cs[ci]; bestF = voiceless ? 0.0 : fr; }
                }
                p.sel_f[f] = bestF;
                p.sel_s[f] = bestS;
            }
        } else {
            if (tid < MAXCAND) {
                int ci = tid + 1;
                double fr = 0.0, st = 0.0, sc = -1e300, lf = -1.0;
                if (ci <= ncand) {
                    fr = S.cf[ci]; st = S.cs[ci];
                    bool voiceless = !(fr > 0.0 && fr < g.ceiling);
                    sc = voiceless ? unvoicedStrength : st - g.octave_cost * log2(g.ceiling / fr);
                    lf = voiceless ? -1.0 : log2(fr);
                }
                size_t o2 = (size_t)f * MAXCAND + tid;
                p.cand_f[o2] = fr; p.cand_s[o2] = st; p.cand_score[o2] = sc; p.cand_lf[o2] = lf;
            }
            if (tid == 0) p.ncand[f] = (uint8_t)ncand;
        }
    }
}

static FrameSmem frames_smem_layout(const PitchPass& p, bool is_cc) {
    int ab = 0, rs = 0, mn = 0, ml = 0;
    for (int k = 0; k < 3; k++) {
        const PitchCfg& g = p.cfg[k];
        int need_a = is_cc ? (int)sizeof(double) * (g.maximumLag + g.nsamp_window + 8) : (int)sizeof(double2) * g.M;
        if (need_a > ab) ab = need_a;
        if (2 * g.brent_ixmax + 1 > rs) rs = 2 * g.brent_ixmax + 1;
        if (g.maxn > mn) mn = g.maxn;
        if (g.maximumLag > ml) ml = g.maximumLag;
    }
    if (mn < MAXCAND) mn = MAXCAND;
    FrameSmem L;
    int o = 0;
    auto take = [&](int bytes) { int r = o; o = (o + bytes + 15) & ~15; return r; };
    L.a = take(ab);
    L.rs = take((int)sizeof(double) * (rs + 1));
    L.pkf = take((int)sizeof(double) * MAXPK);
    L.pks = take((int)sizeof(double) * MAXPK);
    L.pkkey = take((int)sizeof(double) * MAXPK);
    L.cf = take((int)sizeof(double) * (mn + 1));
    L.cs = take((int)sizeof(double) * (mn + 1));
    L.ckey = take((int)sizeof(double) * (mn + 1));
    L.red = take((int)sizeof(double) * 32);
    L.part_stride = (ml + 8) & ~7;
    L.part = take(is_cc ? (int)sizeof(double) * 8 * L.part_stride : 16);
    L.pklag = take((int)sizeof(int) * MAXPK);
    L.cimax = take((int)sizeof(int) * (mn + 1));
    L.masks = take((int)sizeof(unsigned) * 64);
    L.sint = take((int)sizeof(int) * 4);
    L.fi = take(64);
    L.total = o;
    return L;
}

void launch_pitch_frames(const Clips& c, const PitchPass& p, const double2* tw, int max_frames_hint, cudaStream_t s) {
    bool is_cc = p.cfg[0].method != 0;
    FrameSmem L = frames_smem_layout(p, is_cc);
    size_t smem = (size_t)L.total;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int blocks_per_sm = (int)(220 * 1024 / (smem + 1024));
    if (blocks_per_sm > 8) blocks_per_sm = 8;
    if (blocks_per_sm < 1) blocks_per_sm = 1;
    int grid = nsm * blocks_per_sm;
    if (max_frames_hint > 0 && grid > max_frames_hint) grid = max_frames_hint;
    if (grid < 1) grid = 1;
    if (is_cc) {
        cudaFuncSetAttribute(k_pitch_frames<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_pitch_frames<true><<<grid, NTHR, smem, s>>>(c, p, tw, L);
    } else {
        cudaFuncSetAttribute(k_pitch_frames<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_pitch_frames<false><<<grid, NTHR, smem, s>>>(c, p, tw, L);
    }
}

// ------------------------------------------------------------------------------------------------ Viterbi
// fon/Pitch.cpp Pitch_pathFinder: one warp per clip, lane = current candidate, sequential over frames.
__global__ void __launch_bounds__(128) k_pitch_viterbi(Clips c, PitchPass p) {
    const int lane = threadIdx.x & 31;
    const int clip = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (clip >= c.n) return;
    const int nF = p.nF[clip];
    if (nF < 1) return;
    const int f0 = p.fstart[clip];
    const PitchCfg& g = p.cfg[c.cls[clip]];
    const double corr = 0.01 / g.dt;
    const double jumpCost = g.jump_cost * corr, vuvCost = g.vuv_cost * corr;

    double delta = -1e300, lf = -1.0;
    int ncPrev = 0;
    for (int i = 0; i < nF; i++) {
        const size_t fo = (size_t)(f0 + i);
        const int nc = p.ncand[fo];
        double sc = -1e300, clf = -1.0;
        if (lane < MAXCAND) { sc = p.cand_score[fo * MAXCAND + lane]; clf = p.cand_lf[fo * MAXCAND + lane]; }
        if (i == 0) {
            delta = sc; lf = clf; ncPrev = nc;
            continue;
        }
        const bool curVoiceless = clf < 0.0;
        double maximum = -1e30;
        int place = 0;
        for (int c1 = 0; c1 < ncPrev; c1++) {
            double pd = __shfl_sync(FULL_MASK, delta, c1);
            double plf = __shfl_sync(FULL_MASK, lf, c1);
            bool prevVoiceless = plf < 0.0;
            double cost;
            if (curVoiceless) cost = prevVoiceless ? 0.0 : vuvCost;
            else cost = prevVoiceless ? vuvCost : jumpCost * fabs(plf - clf);
            double value = pd - cost + sc;
            if (value > maximum) { maximum = value; place = c1; }
        }
        if (lane < MAXCAND) p.psi[fo * 16 + lane] = (uint8_t)place;
        delta = lane < nc ? maximum : -1e300;
        lf = clf;
        ncPrev = nc;
    }
    // end of the most probable path: first maximum over the last frame's candidates
    int place = 0;
    {
        double best = __shfl_sync(FULL_MASK, delta, 0);
        for (int ci = 1; ci < ncPrev; ci++) {
            double d = __shfl_sync(FULL_MASK, delta, ci);
            if (d > best) { best = d; place = ci; }
        }
    }
    __syncwarp();
    // backtrack in tiles of 32 frames: lanes prefetch the psi rows, then the chain is resolved with shuffles
    for (int hi = nF - 1; hi >= 0; hi -= 32) {
        int lo = hi - 31 < 0 ? 0 : hi - 31;
        int myFrame = hi - lane;                                // lane 0 = frame hi
        unsigned long long r0 = 0, r1 = 0;
        double sf[1];
        if (myFrame >= lo && myFrame >= 1) {
            const unsigned long long* row = (const unsigned long long*)(p.psi + (size_t)(f0 + myFrame) * 16);
            r0 = row[0]; r1 = row[1];
        }
        (void)sf;
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
    k_pitch_viterbi<<<(c.n + 3) / 4, 128, 0, s>>>(c, p);
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
