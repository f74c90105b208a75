// bfe_sortcore.cuh -- pieces shared by the cell-sorted EOF kernels (bfe_sort.cu) and the
// radial-bin-sorted SL kernels (bfe_sl_sort.cu): block scan of a bin histogram, FP64 tensor-core
// MMA, TMA bulk copy + mbarrier helpers.
#pragma once
#include "bfe_device.cuh"
#ifndef BFE_TRACE_PT
#define BFE_TRACE_PT(kid, phase)
#endif

// Slot-claim counters live one per 32-byte L2 sector: same-sector atomics serialise in the L2 (hot cells of a
// disc / cusp sit next to each other in the bin order), ncu + A/B timing in profiles/r01_summary.md section 11.
// They are zeroed by all CTAs of the histogram kernel (a slice each): 8192 one-sector stores from the single
// scanning CTA cost 4 us of serial tail (device timeline, profiles/trace_step.py).
#define BFE_CURSOR_STRIDE 8

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256), 32-byte aligned addresses: one full sector per lane
__device__ __forceinline__ void bfe_st256(void* p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void bfe_ld256_nc(const void* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// exclusive scan of the histogram by one 1024-thread CTA: cell_start = scan; hist cleared.
// Bins are staged through shared memory (s_h, per*1024 ints) with coalesced loads; two levels of warp shuffles.
__device__ __forceinline__ void bfe_block_scan_cells(int ncell, int* __restrict__ hist, int* __restrict__ cell_start,
                                                     int* s_h, int* s_wsum) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (ncell + 1023) / 1024;
    // eight independent L2 loads in flight per thread (one load at a time costs `per` L2 round trips: ncu showed
    // the SM that runs this scan active 9 us longer than the others)
    for (int c0 = tid; c0 < per * 1024; c0 += 8 * 1024) {
        int v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int c = c0 + u * 1024; v[u] = (c < ncell) ? __ldcg(hist + c) : 0; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u * 1024;
            if (c < ncell) hist[c] = 0;
            if (c < per * 1024) s_h[c] = v[u];
        }
    }
    __syncthreads();
    BFE_TRACE_PT(0, 8);
    const int lo = tid * per;
    int sum = 0;
    for (int k = 0; k < per; ++k) sum += s_h[lo + k];
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    BFE_TRACE_PT(0, 9);
    if (warp == 0) {
        int w = s_wsum[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { int v = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += v; }
        s_wsum[lane] = w;
    }
    __syncthreads();
    int run = incl - sum + (warp > 0 ? s_wsum[warp - 1] : 0);    // exclusive prefix of this thread's segment
    BFE_TRACE_PT(0, 10);
    for (int k = 0; k < per; ++k) { int h = s_h[lo + k]; s_h[lo + k] = run; run += h; }
    __syncthreads();
    BFE_TRACE_PT(0, 11);
    for (int c = tid; c < ncell; c += 1024) cell_start[c] = s_h[c];
    if (tid == 1023) cell_start[ncell] = run;
}

__device__ __forceinline__ void bfe_dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---- TMA bulk copy (cp.async.bulk, global -> shared, mbarrier completion) helpers
__device__ __forceinline__ unsigned int bfe_smem_u32(const void* p) {
    return (unsigned int)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bfe_mbar_init(unsigned int bar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bfe_mbar_expect_tx(unsigned int bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bfe_bulk_g2s(unsigned int dst, const void* src, unsigned int bytes, unsigned int bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bfe_mbar_wait(unsigned int bar, unsigned int parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}

