// lds_broadcast_probe.cu -- what does a WARP-UNIFORM 16-byte shared-memory load cost next to a warp-uniform 32-byte
// global load that hits L1?  (Decides whether staging the per-cell / per-interval table blocks in shared memory can
// relieve the L1 data pipe that bounds the coherent per-lane field kernels: ncu l1tex__data_pipe_lsu_wavefronts.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_probe lds_broadcast_probe.cu && ./lds_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k_lds(const double2* __restrict__ g, int iters, int stride, double* out, long long* cyc) {
    extern __shared__ double2 s[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = g[i];
    __syncthreads();
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const int lane_off = (threadIdx.x & 31) * stride;          // stride 0: all lanes the same address (broadcast)
    long long t0 = clock64();
    int idx = (threadIdx.x >> 5) * 8;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const double2 v = s[(idx + u * 4 + lane_off) & 4095];
            a0 += v.x; a1 += v.y;
        }
        idx = (idx + 64) & 4095;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(128) k_ldg(const double2* __restrict__ g, int iters, int stride, double* out, long long* cyc) {
    double a0 = 0, a1 = 0;
    const int lane_off = (threadIdx.x & 31) * stride;
    long long t0 = clock64();
    int idx = (threadIdx.x >> 5) * 8;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double x0, x1, x2, x3;
            const double2* p = g + ((idx + u * 8 + 2 * lane_off) & 4094);
            asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3) : "l"(p));
            a0 += x0 + x2; a1 += x1 + x3;
        }
        idx = (idx + 64) & 4095;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double2* g; double* out; long long* cyc;
    cudaMalloc(&g, 4096 * sizeof(double2)); cudaMemset(g, 0, 4096 * sizeof(double2));
    cudaMalloc(&out, sizeof(double) * sms * 4 * 128); cudaMalloc(&cyc, sizeof(long long) * sms * 4);
    cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 2000;
    for (int ctas = 1; ctas <= 4; ctas *= 2)
        for (int stride = 0; stride <= 1; ++stride) {
            long long h[1024];
            k_lds<<<sms * ctas, 128, 65536>>>(g, iters, stride, out, cyc); cudaDeviceSynchronize();
            k_lds<<<sms * ctas, 128, 65536>>>(g, iters, stride, out, cyc); cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof(long long) * sms * ctas, cudaMemcpyDeviceToHost);
            double m = 0; for (int i = 0; i < sms * ctas; ++i) m += h[i]; m /= sms * ctas;
            // per SM: ctas * 4 warps, each iters*16 LDS.128
            printf("LDS.128  %s  %d CTAs/SM (%2d warps): %.2f cycles per warp-instruction per SM  (%.0f B/clk into registers)\n",
                   stride ? "distinct 16B per lane" : "broadcast            ", ctas, ctas * 4, m / (double)(iters * 16 * ctas * 4),
                   512.0 * iters * 16 * ctas * 4 / m);
            k_ldg<<<sms * ctas, 128>>>(g, iters, stride, out, cyc); cudaDeviceSynchronize();
            k_ldg<<<sms * ctas, 128>>>(g, iters, stride, out, cyc); cudaDeviceSynchronize();
            cudaMemcpy(h, cyc, sizeof(long long) * sms * ctas, cudaMemcpyDeviceToHost);
            m = 0; for (int i = 0; i < sms * ctas; ++i) m += h[i]; m /= sms * ctas;
            printf("LDG.256  %s  %d CTAs/SM (%2d warps): %.2f cycles per warp-instruction per SM  (%.0f B/clk into registers)\n",
                   stride ? "distinct 32B per lane" : "broadcast (L1 hit)   ", ctas, ctas * 4, m / (double)(iters * 8 * ctas * 4),
                   1024.0 * iters * 8 * ctas * 4 / m);
        }
    return 0;
}
