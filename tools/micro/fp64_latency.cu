// Dependent-issue latency of the float64 instructions this pipeline's serial chains are made of, one warp on one SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/micro/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k_lat(double* out, long long* cyc, int iters, double a, double b) {
    double v = threadIdx.x * 1e-3 + 1.0, w = 0.5;
    int idx = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            if (OP == 0) v = fma(v, a, b);
            else if (OP == 1) v = v + a;
            else if (OP == 2) v = v * a;
            else if (OP == 3) { bool p = v > w; w = p ? v : w; v = p ? b + u : v + 1.0; }          // compare + select chain
            else if (OP == 4) v = __shfl_xor_sync(0xffffffffu, v, 1);
            else if (OP == 5) { idx = __shfl_xor_sync(0xffffffffu, idx, 1) + 1; }
            else if (OP == 6) { float f = (float)v; f = f * 1.0001f + 0.5f; v = f; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = v + w + idx;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    const char* names[] = {"DFMA", "DADD", "DMUL", "DSETP+select (v > w ? ...)", "SHFL.BFLY of a double (2 x 32 bit)", "SHFL.BFLY of an int + IADD", "FFMA (f32, with 2 conversions)"};
    const int iters = 2048;
    for (int op = 0; op < 7; op++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (op) {
            case 0: k_lat<0><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); break;
            case 1: k_lat<1><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); break;
            case 2: k_lat<2><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); break;
            case 3: k_lat<3><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); break;
            case 4: k_lat<4><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); break;
            case 5: k_lat<5><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); break;
            case 6: k_lat<6><<<1, 32>>>(out, cyc, iters, 0.999999, 1e-9); break;
            }
            cudaDeviceSynchronize();
        }
        long long c = 0;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-40s %7.1f cycles per dependent step\n", names[op], (double)c / (iters * 16.0));
    }
    return 0;
}
