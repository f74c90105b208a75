// Issue rate of the FP64 pipe for DFMA, DMUL, DADD and the field kernels' mix (835 : 577 : 142), alone and with a stream of
// L1-resident 256-bit loads beside it (the per-point field kernels issue 84 such loads per 1554 FP64 instructions).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_mix profiles/probes/fp64_mix_probe.cu && /tmp/fp64_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int LOADS, int NCH = 8>   // NCH independent chains per thread;   // KIND 0 dfma, 1 dmul, 2 dadd, 3 mix (3 fma : 2 mul : ~0.5 add)
__global__ void __launch_bounds__(128) probe(double* out, const double4* __restrict__ tab, int iters, long long* cyc) {
    double a[NCH];
    for (int k = 0; k < NCH; ++k) a[k] = 1.0 + 1e-9 * (threadIdx.x + k);
    const double b = 1.0 + 1e-12 * threadIdx.x, c = 1e-13 * blockIdx.x;
    double4 acc = make_double4(0, 0, 0, 0);
    const double4* tp = tab + (threadIdx.x & 31) * 2;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if (KIND == 0) a[k] = fma(a[k], b, c);
                else if (KIND == 1) a[k] = __dmul_rn(a[k], b);
                else if (KIND == 2) a[k] = __dadd_rn(a[k], c);
                else {
                    const int sel = (r * NCH + k) % 11;          // 6 fma, 4 mul, 1 add per 11
                    if (sel < 6) a[k] = fma(a[k], b, c);
                    else if (sel < 10) a[k] = __dmul_rn(a[k], b);
                    else a[k] = __dadd_rn(a[k], c);
                }
            }
        }
        if (LOADS) {
#pragma unroll
            for (int l = 0; l < LOADS; ++l) {
                double4 v;
                asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(tp + 64 * ((it + l) & 7)));
                acc.x += v.x;
            }
        }
    }
    const long long t1 = clock64();
    double s = acc.x;
    for (int k = 0; k < NCH; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int KIND, int LOADS, int NCH = 8>
static void run(const char* name, int ctas_per_sm, double* out, const double4* tab, long long* cyc) {
    const int iters = 2000;
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    probe<KIND, LOADS, NCH><<<sms * ctas_per_sm, 128>>>(out, tab, 10, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<KIND, LOADS, NCH><<<sms * ctas_per_sm, 128>>>(out, tab, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double winst = (double)iters * 4 * NCH * 4 * ctas_per_sm;      // FP64 warp instructions per SM
    // whole-grid rate from the event time (the oldest CTA of an SM is served first -- greedy scheduling -- so the cycle count of
    // block 0 is the time of ONE CTA running alone, printed for reference)
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    printf("%-22s chains %d ctas/SM %d  warps/SM %2d  loads/iter %d : %.3f FP64 warp-instr / clk / SM (peak 2)  (%.1f us; block 0 alone: %lld clk)\n",
           name, NCH, ctas_per_sm, 4 * ctas_per_sm, LOADS, winst / ((double)ms * 1e-3 * khz * 1e3), ms * 1e3, c);
}

int main() {
    double* out; double4* tab; long long* cyc;
    cudaMalloc(&out, sizeof(double) * 148 * 16 * 128);
    cudaMalloc(&tab, sizeof(double4) * 4096);
    cudaMemset(tab, 0, sizeof(double4) * 4096);
    cudaMalloc(&cyc, sizeof(long long));
    for (int c : {3, 4, 8}) {
        run<0, 0>("DFMA", c, out, tab, cyc);
        run<1, 0>("DMUL", c, out, tab, cyc);
        run<2, 0>("DADD", c, out, tab, cyc);
        run<3, 0>("mix 6:4:1", c, out, tab, cyc);
        run<3, 2>("mix + 2 LDG.256/32", c, out, tab, cyc);     // 2 loads per 32 FP64: the field kernels' ratio is 84 / 1554 = 1.7 per 32
        run<3, 4>("mix + 4 LDG.256/32", c, out, tab, cyc);
        run<0, 2>("DFMA + 2 LDG.256/32", c, out, tab, cyc);
        run<3, 0, 2>("mix 6:4:1", c, out, tab, cyc);
        run<3, 2, 2>("mix + 2 LDG.256/32", c, out, tab, cyc);
        run<3, 0, 4>("mix 6:4:1", c, out, tab, cyc);
    }
    return 0;
}
