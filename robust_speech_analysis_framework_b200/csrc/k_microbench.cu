// k_microbench.cu -- measured denominators for the compute-side roofline of this float64 pipeline.
//
// MEASURED_PEAKS.json (driver-written) holds an HBM copy bandwidth and a bf16 tensor-core figure; neither bounds the MSHDS
// path, whose arithmetic is float64 on the vector pipe (~1e4 FLOP per compulsory byte).  bench.py therefore
// measures the DFMA issue peak of the device it runs on, in the same process as the timed steps, and reports the pipeline's
// float64 FLOP rate against it.  No reference counterpart (the reference has no GPU path).
#include "../../include/mshds_b200.h"
#include "internal.h"

// CHAINS independent dependent-FMA chains per thread: enough ILP to cover the DFMA latency at 8 warps per scheduler.
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double v[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; k++) v[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < CHAINS; k++) v[k] = fma(v[k], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < CHAINS; k++) s += v[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

// Returns the best of `reps` timed launches in TFLOP/s (2 FLOP per DFMA); < 0 on a CUDA error.
double run_dfma_peak(double* scratch, cudaStream_t s, int reps) {
    const int nsm = sm_count();
    const int grid = nsm * 8, block = 256, iters = 4096, chains = 8;
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.0;
    double best = -1.0;
    for (int r = 0; r < reps + 1; r++) {          // first launch is the warm-up
        cudaEventRecord(e0, s);
        k_dfma_peak<chains><<<grid, block, 0, s>>>(scratch, iters, 0.999999, 1e-9);
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * (double)grid * block * chains * 8.0 * iters;
        const double tf = flop / ((double)ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}
