// stage.cuh -- sample spans staged in shared memory by the TMA engine (cp.async.bulk, 1-D bulk tensor copy) with mbarrier
// completion, double buffered: the frame kernels (k_acw.cu, k_moments.cu, k_lld.cu) work on one span while thread 0 already
// fetches the next.  A span is the contiguous run of samples that the <= 8 consecutive frames of one CTA turn touch; the
// recordings stay int16 in HBM (2 B/sample, or float64 behind the resampling front-end), are fetched once per turn and
// converted on read.
//
// Alignment: bulk copies move whole 16-byte units, so the source is aligned down / up and `shift` says where element 0
// landed; the few elements of a 16-byte unit that would reach outside the chunk's sample array are left out of the bulk copy
// and filled in by plain loads (head / tail0).
#pragma once
#include "common.cuh"

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar` (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init_pair(unsigned long long* bars) {      // thread 0, before the first __syncthreads()
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

struct StageSeg {             // one staged segment: consecutive frames of ONE clip (written by thread 0, read by all)
    int f0, n;                // first flat frame index, frames (0 = no more work)
    int clip, cls, k0;        // clip, speaker class, index of the first frame inside the clip
    long long base;           // chunk offset of the clip's first sample
    long long sA;             // 1-based clip sample index of staged element 0
    int count;                // staged elements
    int shift;                // staged element e sits `shift + e` elements into the stage buffer (16-byte alignment of the source)
    int head, tail0;          // elements [0, head) and [tail0, count) are NOT covered by the bulk copy: copied by threads
    int tma_bytes;            // 0 = nothing was issued (barrier not armed)
};

struct StageParams {
    const unsigned char* pcm_bytes;   // chunk samples as bytes
    long long total_elems;            // samples in the chunk
    int esz;                          // 2 (int16) or 8 (float64 behind the resampling front-end)
    int stage_bytes;                  // bytes of ONE stage buffer
};

__device__ __forceinline__ double staged(const unsigned char* st, int esz, int e /* element index incl. shift */) {
    return esz == 2 ? (double)((const short*)st)[e] * (1.0 / 32768.0) : ((const double*)st)[e];
}

// thread 0: bulk copy of clip samples [sA, sA + count) (1-based, already clamped to the clip) into `stage`, armed on `bar`
__device__ __forceinline__ void stage_issue(const StageParams& A, StageSeg* sg, long long base, long long sA, int count,
                                            unsigned char* stage, unsigned long long* bar) {
    sg->base = base; sg->sA = sA; sg->count = count;
    const int esz = A.esz;
    const long long gA = base + sA - 1;                                   // chunk element index of staged element 0
    const unsigned long long addr = (unsigned long long)(A.pcm_bytes + gA * esz);
    const unsigned long long a0 = addr & ~15ull;
    sg->shift = (int)((addr - a0) / esz);
    const unsigned long long a1 = (addr + (unsigned long long)count * esz + 15ull) & ~15ull;
    // bytes the bulk copy may touch: whole 16-byte units inside the chunk's sample array
    const unsigned long long lo = ((unsigned long long)A.pcm_bytes + 15ull) & ~15ull;
    const unsigned long long hi = ((unsigned long long)A.pcm_bytes + (unsigned long long)A.total_elems * esz) & ~15ull;
    unsigned long long t0 = a0 > lo ? a0 : lo, t1 = a1 < hi ? a1 : hi;
    if (t1 > a0 + (unsigned long long)A.stage_bytes) t1 = a0 + (unsigned long long)A.stage_bytes;    // (cannot happen: stage sized for the worst span)
    int head = count, tail0 = count, bytes = 0;                           // default: everything by threads
    if (count > 0 && t1 > t0) {
        bytes = (int)(t1 - t0);
        head = t0 > addr ? (int)((t0 - addr + esz - 1) / esz) : 0;
        tail0 = (int)((t1 - addr) / esz);
        if (tail0 > count) tail0 = count;
        mbar_expect_tx(bar, (unsigned)bytes);
        tma_load_1d(stage + (t0 - a0), (const void*)t0, (unsigned)bytes, bar);
    }
    sg->head = head; sg->tail0 = tail0; sg->tma_bytes = bytes;
}

// all threads of the CTA: elements the bulk copy could not cover (unaligned ends of the chunk), then wait for the copy
template <int NT>
__device__ __forceinline__ void stage_complete(const StageParams& A, const StageSeg& sg, unsigned char* st, unsigned long long* bar,
                                               unsigned& phase) {
    if (sg.head > 0 || sg.tail0 < sg.count) {
        const long long g0 = sg.base + sg.sA - 1;
        for (int e = threadIdx.x; e < sg.count; e += NT) {
            if (e >= sg.head && e < sg.tail0) continue;
            if (A.esz == 2) ((short*)st)[sg.shift + e] = ((const short*)A.pcm_bytes)[g0 + e];
            else ((double*)st)[sg.shift + e] = ((const double*)A.pcm_bytes)[g0 + e];
        }
        __syncthreads();
    }
    if (sg.tma_bytes > 0) {
        mbar_wait(bar, phase);
        phase ^= 1;
    }
}
