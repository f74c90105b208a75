// bfe_peak.cu -- measured FP64 peaks of the device the library runs on: the denominators of the FP64 side of the
// roofline ("report against the MEASURED DFMA peak").
//
//   kind 0 : vector FP64 -- eight independent DFMA chains per thread, 1024 threads per SM resident 2x over
//   kind 1 : tensor FP64 -- mma.sync.m8n8k4.f64 (SASS DMMA), four independent accumulator pairs per warp
//
// Both loops are pure register arithmetic (no memory traffic); the result of every chain is folded into one store per
// thread so the compiler cannot drop the work.  Timed with CUDA events on the caller's stream, best of `reps` launches.
#include "bfe_sortcore.cuh"

__global__ void __launch_bounds__(256)
fp64_dfma_peak_kernel(int iters, double a, double b, double* __restrict__ out) {
    double c0 = threadIdx.x * 1e-9, c1 = c0 + 1.0, c2 = c0 + 2.0, c3 = c0 + 3.0, c4 = c0 + 4.0, c5 = c0 + 5.0, c6 = c0 + 6.0, c7 = c0 + 7.0;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
            c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
}

__global__ void __launch_bounds__(256)
fp64_dmma_peak_kernel(int iters, double a, double b, double* __restrict__ out) {
    double c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) c[k] = threadIdx.x * 1e-9 + k;
    const double av = a + 1e-12 * (threadIdx.x & 31), bv = b - 1e-12 * (threadIdx.x & 31);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bfe_dmma_m8n8k4(c[0], c[1], av, bv); bfe_dmma_m8n8k4(c[2], c[3], av, bv);
            bfe_dmma_m8n8k4(c[4], c[5], av, bv); bfe_dmma_m8n8k4(c[6], c[7], av, bv);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
}

extern "C" int bfe_fp64_peak(int kind, double* tflops, void* stream_) {
    if (!tflops || kind < 0 || kind > 1) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    BFE_CUDA(cudaGetDevice(&dev));
    BFE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 8, block = 256;                       // 2048 threads per SM: every scheduler has 16 warps to pick from
    const int iters = kind == 0 ? 4096 : 2048;
    double* out = nullptr;
    BFE_CUDA(cudaMalloc(&out, sizeof(double) * (size_t)grid * block));
    cudaEvent_t e0, e1;
    BFE_CUDA(cudaEventCreate(&e0)); BFE_CUDA(cudaEventCreate(&e1));
    // flop per launch: DFMA = 2 per lane per instruction; DMMA m8n8k4 = 8*8*4 FMA = 512 flop per warp per instruction
    const double flop = kind == 0 ? (double)grid * block * (double)iters * 64.0 * 2.0
                                  : (double)grid * (block / 32) * (double)iters * 16.0 * 512.0;
    double best = 0.0;
    for (int r = 0; r < 4; ++r) {
        BFE_CUDA(cudaEventRecord(e0, stream));
        if (kind == 0) fp64_dfma_peak_kernel<<<grid, block, 0, stream>>>(iters, 0.999999, 1e-7, out);
        else fp64_dmma_peak_kernel<<<grid, block, 0, stream>>>(iters, 0.5, 0.25, out);
        BFE_LAUNCH_CHECK("fp64_peak_kernel");
        BFE_CUDA(cudaEventRecord(e1, stream));
        BFE_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        BFE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms > 0.f) { const double t = flop / (ms * 1e-3) / 1e12; if (t > best) best = t; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return BFE_OK;
}
