// fftwarp.cuh -- warp-level pieces of the register-resident packed real FFT (see fftreg.cuh for the decomposition):
// the two-pass transform with one shared-memory exchange, and the real-input untangle done with warp shuffles.
// Used by the autocorrelation pitch frames (k_acw.cu), the spectrogram moments (k_moments.cu), the power cepstrogram
// (k_cpp.cu) and the MFCC frames (k_lld.cu).  L = M / 32 lanes own one frame (L = 32: a warp, L = 16: half a warp).
#pragma once
#include "common.cuh"
#include "fftreg.cuh"

// Pass twiddles exp(-2 pi i j q / M), row q of a [32][L] table, as seen by lane j.  FwTwGlobal reads the table through the
// read-only cache; FwTwShared reads a copy the CTA staged in shared memory (fw_stage_twiddles) -- the frame kernels run only two
// warps per scheduler (255 registers), which cannot hide an L1/L2 round trip in front of every twiddle multiply (ncu, round 2:
// 40 % of the stall samples of k_ac_frames_w were DMULs waiting on the long scoreboard), but do hide a shared-memory one.
struct FwTwGlobal {
    const double2* base; int L;                 // base = table + j
    __device__ __forceinline__ FwTwGlobal(const double2* twb, int j, int L_) : base(twb + j), L(L_) {}
    __device__ __forceinline__ double2 operator()(int q) const { return __ldg(base + q * L); }
};
struct FwTwShared {
    const double2* base; int L;                 // base = shared-memory copy of the table + j
    __device__ __forceinline__ FwTwShared(const double2* twb_smem, int j, int L_) : base(twb_smem + j), L(L_) {}
    __device__ __forceinline__ double2 operator()(int q) const { return base[q * L]; }
};
// all NT threads of the CTA: copy a [32][L] table (32 L complex doubles) into shared memory; the caller synchronises
template <int NT>
__device__ __forceinline__ void fw_stage_twiddles(double2* dst_smem, const double2* __restrict__ twb, int L) {
    for (int i = threadIdx.x; i < 32 * L; i += NT) dst_smem[i] = __ldg(twb + i);
}

// One transform of the frame(s) of this warp.  `a` holds pass-A input in logical order (element k = z[j + L k]); on return
// `a` holds logical element r (k = j + L r) of the result at a[fr_slot<L>(r)].  xch: this lane's frame exchange region
// (M complex doubles); twb: [32][L] table of exp(-2 pi i j q / M).
template <int L, int SIGN, class TW>
__device__ __forceinline__ void fw_transform(double2 (&a)[32], double2* xch, int j, const TW twf) {
    fr_fft<32, SIGN>(a);
    fr_static_for<0, 32>([&](auto qc) {
        constexpr int q = decltype(qc)::value;
        double2 v = a[fr_brev(q, 5)];
        if constexpr (q > 0) {
            double2 t = twf(q);                                 // exp(-2 pi i j q / M); conjugate for the inverse
            if (SIGN > 0) t.y = -t.y;
            v = fr_mul(v, t);
        }
        xch[q * L + (j ^ (q & 7))] = v;
    });
    __syncwarp();
    if constexpr (L == 32) {
        const int q = j;
        fr_static_for<0, 32>([&](auto jc) {
            constexpr int jj = decltype(jc)::value;
            a[jj] = xch[q * 32 + (jj ^ (q & 7))];
        });
        fr_fft<32, SIGN>(a);
    } else {
        fr_static_for<0, 32>([&](auto sc) {
            constexpr int sl = decltype(sc)::value, h = sl >> 4, jj = sl & 15;
            const int q = j + 16 * h;
            a[sl] = xch[q * 16 + (jj ^ (q & 7))];
        });
        fr_fft<16, SIGN>(&a[0]);
        fr_fft<16, SIGN>(&a[16]);
    }
    __syncwarp();
}

__device__ __forceinline__ double2 fw_shfl2(double2 v, int src) {
    return make_double2(__shfl_sync(FULL_MASK, v.x, src), __shfl_sync(FULL_MASK, v.y, src));
}

struct FwIdentity { __device__ __forceinline__ double operator()(double p) const { return p; } };

// Power spectrum of the packed real transform, in place: Z (slots) -> Y (slots), the packed input of the inverse transform
// of F(|X[k]|^2) (F = identity: autocorrelation; F = log: cepstrum).  Bin k = j + L r pairs with M - k = (L - j) + L (31 - r),
// held by lane L - j: every lane computes the pairs of its r < 16 (both members) and trades the results by shuffle.  Lane
// j = 0 pairs inside itself (r with 32 - r; bins 0 and M/2 are their own partners).  wj = exp(-2 pi i j / N), N = 2M.
template <int L, class F>
__device__ __forceinline__ void fw_power_retangle(double2 (&e)[32], int lane, int j, double2 wj, F f) {
    const int grp = lane & ~(L - 1);
    const int pl = grp | ((L - j) & (L - 1));
    const bool j0 = j == 0;
    fr_static_for<0, 16>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        const double2 mine = e[fr_slot<L>(r)];
        double2 theirs = fw_shfl2(e[fr_slot<L>(31 - r)], pl);
        if constexpr (r >= 1) { if (j0) theirs = e[fr_slot<L>(32 - r)]; }
        const double2 wk = fr_mul(wj, make_double2(fr_cos64(r), -fr_sin64(r)));      // exp(-2 pi i (j + L r) / N)
        double pk, pmk;
        fr_pair_powers(mine, theirs, wk, &pk, &pmk);
        double2 yk, ymk;
        fr_pair_retangle(f(pk), f(pmk), wk, &yk, &ymk);
        if constexpr (r == 0) {
            if (j0) {
                const double p0 = f((mine.x + mine.y) * (mine.x + mine.y)), pM = f((mine.x - mine.y) * (mine.x - mine.y));
                yk = make_double2(p0 + pM, p0 - pM);
            }
        }
        const double2 ret = fw_shfl2(ymk, pl);
        e[fr_slot<L>(r)] = yk;
        if (!j0) e[fr_slot<L>(31 - r)] = ret;
        if constexpr (r >= 1) { if (j0) e[fr_slot<L>(32 - r)] = ymk; }
    });
    {   // lane j = 0: bin M/2 pairs with itself
        const double2 z = e[fr_slot<L>(16)];
        const double2 wk = make_double2(fr_cos64(16), -fr_sin64(16));
        double pk, pmk;
        fr_pair_powers(z, z, wk, &pk, &pmk);
        double2 yk, ymk;
        fr_pair_retangle(f(pk), f(pmk), wk, &yk, &ymk);
        if (j0) e[fr_slot<L>(16)] = yk;
    }
}

// Powers only: P[r] = |X[j + L r]|^2 for r < RMAX (RMAX <= 32) of the real transform whose packed spectrum sits in e (slots).
// (The bin M itself -- Nyquist -- is not produced.)
template <int L, int RMAX>
__device__ __forceinline__ void fw_powers(const double2 (&e)[32], int lane, int j, double2 wj, double (&P)[32]) {
    const int grp = lane & ~(L - 1);
    const int pl = grp | ((L - j) & (L - 1));
    const bool j0 = j == 0;
    fr_static_for<0, 16>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        // the pair (r, 31 - r) yields P[r] here and, through the partner lane, P[31 - r] there: skip pairs nobody needs
        if constexpr (r < RMAX || 31 - r < RMAX || 32 - r < RMAX) {
            const double2 mine = e[fr_slot<L>(r)];
            double2 theirs = fw_shfl2(e[fr_slot<L>(31 - r)], pl);
            if constexpr (r >= 1) { if (j0) theirs = e[fr_slot<L>(32 - r)]; }
            const double2 wk = fr_mul(wj, make_double2(fr_cos64(r), -fr_sin64(r)));
            double pk, pmk;
            fr_pair_powers(mine, theirs, wk, &pk, &pmk);
            if constexpr (r == 0) { if (j0) pk = (mine.x + mine.y) * (mine.x + mine.y); }
            if constexpr (r < RMAX) P[r] = pk;
            if constexpr (31 - r < RMAX) {
                const double got = __shfl_sync(FULL_MASK, pmk, pl);
                if (!j0) P[31 - r] = got;
            }
            if constexpr (r >= 1 && 32 - r < RMAX) { if (j0) P[32 - r] = pmk; }
        }
    });
    if constexpr (16 < RMAX) {
        const double2 z = e[fr_slot<L>(16)];
        double pk, pmk;
        fr_pair_powers(z, z, make_double2(fr_cos64(16), -fr_sin64(16)), &pk, &pmk);
        if (j0) P[16] = pk;
    }
}

// ------------------------------------------------------------------------------------------------ looped round trip
// forward packed real FFT -> G(|X|^2) -> inverse FFT with ONE copy of the transform code, for a run-time L (16 or 32 lanes per
// frame).  Both sizes share the code because fr_slot<16> == fr_slot<32> == the 5-bit reversal and pass B of the 512-point case
// is the 32-point butterfly network minus its first stage (= two 16-point transforms); forward and inverse share it because
// inverse = conj FFT conj, the conjugations folded into the re-tangle step here and into the caller's read-out.
//   in : a[k] = z[j + L k] (logical order), z[n] = x[2n] + i x[2n+1] the packed real frame
//   out: a[fr_brev(r, 5)] = conj(y[j + L r]),  y[n] = c[2n] + i c[2n+1],  c = sum_{k<N} G(|X[k]|^2) exp(+2 pi i k n / N)
// (G = identity: autocorrelation of the frame; G = log: its real cepstrum).  Keeping the unrolled code small matters: the
// first warp-per-frame kernel of this round had four differently unrolled transforms per size and spent 35 % of its stall
// samples waiting for instruction fetch (ncu stalled_no_instructions; the L1.5 instruction cache holds 32 KB).
__device__ __forceinline__ void fw_fft32(double2 (&a)[32], bool two_halves) {
    if (!two_halves) FrDifStage<32, 16, -1, 0>::run(a);      // first stage pairs (i, i + 16); without it: two 16-point FFTs
    fr_dif_stages<32, 8, -1>(a);
}

template <class G, class TW>
__device__ __forceinline__ void fw_roundtrip(double2 (&a)[32], double2* xch, int lane, int j, int L, const TW twf,
                                             double2 wj /* exp(-2 pi i j / N) */, G gf) {
    const int grp = lane & ~(L - 1);
    const int pl = grp | ((L - j) & (L - 1));
    const bool j0 = j == 0;
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
        // pass A: 32-point transform of the lane's stride-L elements, twiddle, transpose through shared memory
        fw_fft32(a, false);
        fr_static_for<0, 32>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            double2 v = a[fr_brev(q, 5)];
            if constexpr (q > 0) v = fr_mul(v, twf(q));                   // exp(-2 pi i j q / M)
            xch[q * L + (j ^ (q & 7))] = v;
        });
        __syncwarp();
        fr_static_for<0, 32>([&](auto sc) {
            constexpr int sl = decltype(sc)::value;
            const int q = L == 32 ? j : j + 16 * (sl >> 4), jj = L == 32 ? sl : (sl & 15);
            a[sl] = xch[q * L + (jj ^ (q & 7))];
        });
        // pass B: rows of L-point transforms (L = 16: two rows per lane = the 32-point network without its first stage)
        fw_fft32(a, L == 16);
        __syncwarp();
        if (it == 0) {
            // a[brev5(r)] = Z[j + L r]: powers, G, re-tangled into the CONJUGATE of the packed inverse-transform input; pairs are
            // traded with lane L - j by shuffle (see fw_power_retangle)
            fr_static_for<0, 16>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                const double2 mine = a[fr_brev(r, 5)];
                double2 theirs = fw_shfl2(a[fr_brev(31 - r, 5)], pl);
                if constexpr (r >= 1) { if (j0) theirs = a[fr_brev(32 - r, 5)]; }
                const double2 wk = fr_mul(wj, make_double2(fr_cos64(r), -fr_sin64(r)));      // exp(-2 pi i (j + L r) / N)
                double pk, pmk;
                fr_pair_powers(mine, theirs, wk, &pk, &pmk);
                double2 yk, ymk;
                fr_pair_retangle(gf(pk), gf(pmk), wk, &yk, &ymk);
                if constexpr (r == 0) {
                    if (j0) {
                        const double p0 = gf((mine.x + mine.y) * (mine.x + mine.y)), pM = gf((mine.x - mine.y) * (mine.x - mine.y));
                        yk = make_double2(p0 + pM, p0 - pM);
                    }
                }
                yk.y = -yk.y; ymk.y = -ymk.y;
                const double2 ret = fw_shfl2(ymk, pl);
                a[fr_brev(r, 5)] = yk;
                if (!j0) a[fr_brev(31 - r, 5)] = ret;
                if constexpr (r >= 1) { if (j0) a[fr_brev(32 - r, 5)] = ymk; }
            });
            {   // lane j = 0: bin M/2 pairs with itself
                const double2 z = a[fr_brev(16, 5)];
                const double2 wk = make_double2(fr_cos64(16), -fr_sin64(16));
                double pk, pmk;
                fr_pair_powers(z, z, wk, &pk, &pmk);
                double2 yk, ymk;
                fr_pair_retangle(gf(pk), gf(pmk), wk, &yk, &ymk);
                yk.y = -yk.y;
                if (j0) a[fr_brev(16, 5)] = yk;
            }
            // slots -> logical order for the next pass A: the 5-bit reversal is an involution, 12 swaps
            fr_static_for<0, 32>([&](auto rc) {
                constexpr int r = decltype(rc)::value, br = fr_brev(r, 5);
                if constexpr (r < br) { const double2 tmp = a[r]; a[r] = a[br]; a[br] = tmp; }
            });
        }
    }
}
