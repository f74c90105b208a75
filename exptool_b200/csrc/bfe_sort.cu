// bfe_sort.cu -- cell-sorted formulations of the EOF passes.
//
// The direct kernels (bfe_eof.cu) gather four table rows per particle: 7.5 kB of L2
// traffic per 32 B of particle data, which binds on the L1/L2 path (ncu, profiles/).
// Bilinear interpolation is linear in the table, so for all particles that share a cell
//
//   sum_p m_p trig_ch(p) interp_p(T[j])  =  sum_{k in 4 corners} T[node_k][j] * S_k,ch
//   S_k,ch = sum_{p in cell} c_k(p) m_p trig_ch(p)                     (52 sums per cell)
//
// ("deposit then contract").  Particles are counting-sorted by cell
// (integer keys, integer cursors), the 52 sums are formed per run of equal cells on the
// FP64 tensor cores, and the table rows are read once per CELL instead of once per
// particle.  No floating-point atomics anywhere.
//
//   eof_cell_hist_kernel     : per-particle cell id -> histogram (shared-memory int counters);
//                              the last CTA to finish scans it (cell_start, cursors)
//   eof_cell_scatter_kernel  : 64-byte record {c00,c10,c01,c11,cos phi,sin phi,m|R,cell:perm}
//                              written at its sorted position
//   eof_segsum_kernel        : runs -> S (one 8x8 FP64 tensor-core tile per run segment, stored to global)
//   eof_node_contract_kernel : S, table rows -> coefficient partials -> last-CTA reduce
//   eof_force_sorted_mma_kernel : field evaluation in sorted order on the FP64 tensor cores
//   eof_force_sorted_kernel  : the same per lane (warp-uniform table rows; option force_mma = 0),
//                              outputs scattered back to the caller's particle order
#include "bfe_device.cuh"

// Optional device-side timeline (build with BFE_NVCC_FLAGS=-DBFE_TRACE; profiles/trace_step.py): thread 0 of
// every CTA logs %globaltimer at the phase boundaries of the step's kernels.
#ifdef BFE_TRACE
__device__ unsigned long long* g_trace = nullptr;
__device__ unsigned int g_trace_cap = 0;
extern "C" int bfe_debug_set_trace(unsigned long long* p, unsigned int cap) {
    cudaError_t e = cudaMemcpyToSymbol(g_trace, &p, sizeof(p));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap));
    return (int)e;
}
__device__ __forceinline__ void bfe_trace(int kid, int phase) {
    if (threadIdx.x == 0 && g_trace) {
        unsigned long long t; unsigned int sm;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        const unsigned long long slot = atomicAdd(g_trace, 1ull);
        if (slot < g_trace_cap) {
            g_trace[2 + 2 * slot] = ((unsigned long long)kid << 56) | ((unsigned long long)phase << 48) |
                                    ((unsigned long long)sm << 32) | (unsigned long long)blockIdx.x;
            g_trace[3 + 2 * slot] = t;
        }
    }
}
#define BFE_TRACE_PT(kid, phase) bfe_trace(kid, phase)
#else
#define BFE_TRACE_PT(kid, phase)
#endif

#include "bfe_sortcore.cuh"

struct __align__(16) EofRec {
    double c00, c10, c01, c11;   // bilinear weights (eof.py:448-451), NOT mass weighted
    double c1, s1;               // cos phi, sin phi
    double aux;                  // mass (0 if the set was prepared without masses)
    unsigned long long cellperm; // (cell << 32) | original particle index
};

__device__ __forceinline__ int bfe_cell_of(const EofGeom& g, const EofBin& b) {
    // b.node = ix*ny1 + iy  ->  cell = ix*numy + iy
    int ix = b.node / g.ny1;
    int iy = b.node - ix * g.ny1;
    return ix * g.numy + iy;
}

// histogram of cell ids (shared-memory int counters per CTA, merged with one global int add per
// non-empty bin); the last CTA to finish scans the histogram, so no separate scan launch is needed.
// dynamic smem: per*1024 ints (>= ncell).
__global__ void __launch_bounds__(1024)
eof_cell_hist_kernel(EofGeom g, int ncell, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                     const double* __restrict__ z, int* __restrict__ hist, int* __restrict__ cell_start,
                     int* __restrict__ cursor, unsigned int* __restrict__ counter, int* __restrict__ cellid) {
    extern __shared__ int s_hist[];
    __shared__ int s_wsum[32];
    __shared__ bool s_last;
    BFE_TRACE_PT(0, 0);
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) s_hist[c] = 0;
    bfe_pdl_wait();                               // everything above overlaps the previous kernel's tail
    bfe_pdl_trigger();
    // this CTA's slice of the slot-claim counters (used by the scatter kernel) is cleared here, in parallel
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x)
        cursor[(size_t)c * BFE_CURSOR_STRIDE] = 0;
    __syncthreads();
    BFE_TRACE_PT(0, 1);
    // two particles per thread per pass: six independent loads in flight
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 2 * stride) {
        const int64_t i2 = i + stride;
        const bool two = i2 < n;
        const double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        const double qx = two ? __ldg(x + i2) : 1.0, qy = two ? __ldg(y + i2) : 0.0, qz = two ? __ldg(z + i2) : 0.0;
        int cell, cell2;
        if (!bfe_eof_cell_fast(g, px, py, pz, cell)) {            // near a cell edge / unusual input: exact FP64 index
            double r = sqrt(px * px + py * py + 1.e-10);
            cell = bfe_eof_bin(g, r, pz).cell;
        }
        atomicAdd(&s_hist[cell], 1);
        cellid[i] = cell;                 // kept for the scatter kernel (the inverse-permutation array, overwritten there)
        if (two) {
            if (!bfe_eof_cell_fast(g, qx, qy, qz, cell2)) {
                double r = sqrt(qx * qx + qy * qy + 1.e-10);
                cell2 = bfe_eof_bin(g, r, qz).cell;
            }
            atomicAdd(&s_hist[cell2], 1);
            cellid[i2] = cell2;
        }
    }
    __syncthreads();
    BFE_TRACE_PT(0, 2);
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
        int v = s_hist[c];
        if (v) atomicAdd(&hist[c], v);
    }
    BFE_TRACE_PT(0, 3);
    __threadfence();
    __syncthreads();
    BFE_TRACE_PT(0, 4);
    if (threadIdx.x == 0) {
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    BFE_TRACE_PT(0, 5);
    if (s_last) {
        __threadfence();
        BFE_TRACE_PT(0, 6);
        bfe_block_scan_cells(ncell, hist, cell_start, s_hist, s_wsum);
        if (threadIdx.x == 0) *counter = 0u;
        __syncthreads();
        BFE_TRACE_PT(0, 7);
    }
}

__global__ void __launch_bounds__(256)
eof_cell_scatter_kernel(EofGeom g, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                        const double* __restrict__ z, const double* __restrict__ mass,
                        const int* __restrict__ cell_start, int* __restrict__ cursor, EofRec* __restrict__ rec,
                        int* __restrict__ inv, double* __restrict__ r_orig) {
    // Two particles per thread per pass.  Order of work per pass: (1) all eight loads, (2) the cell id the histogram
    // kernel left in inv[] (FP32 fast path, exact by construction, FP64 fallback near edges), (3) the integer slot claims, (4) the FP64 bin
    // fractions, weights and cos/sin phi WHILE the claims are in flight (they were 1/3 of this kernel's stall
    // samples when issued after the FP64 arithmetic, ncu profiles/), (5) the record stores.
    constexpr int U = 2;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    BFE_TRACE_PT(1, 0);
    for (int64_t base = (int64_t)blockIdx.x * (256 * U); base < n; base += (int64_t)gridDim.x * (256 * U)) {
        double px[U], py[U], pz[U], aux[U];
        int64_t idx[U];
        int cell[U], pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            idx[u] = base + u * 256 + threadIdx.x;
            const bool on = idx[u] < n;
            px[u] = on ? __ldg(x + idx[u]) : 1.0; py[u] = on ? __ldg(y + idx[u]) : 0.0; pz[u] = on ? __ldg(z + idx[u]) : 0.0;
            aux[u] = (on && mass) ? __ldg(mass + idx[u]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cell[u] = (idx[u] < n) ? inv[idx[u]] : 0;   // cell id left here by the histogram kernel
#pragma unroll
        for (int u = 0; u < U; ++u)
            pos[u] = (idx[u] < n) ? (__ldg(cell_start + cell[u]) + atomicAdd(&cursor[(size_t)cell[u] * BFE_CURSOR_STRIDE], 1)) : 0;   // integer slot claim, not a data reduction
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double r = sqrt(px[u] * px[u] + py[u] * py[u] + 1.e-10);   // eof.py:531 / 1070
            const EofBin b = bfe_eof_bin(g, r, pz[u]);                        // b.cell == cell[u]
            double c1, s1;
            bfe_cossin_phi(px[u], py[u], c1, s1);
            if (idx[u] < n) {
                unsigned long long cp = ((unsigned long long)(unsigned int)cell[u] << 32) |
                                        (unsigned long long)(unsigned int)idx[u];
                char* dst = reinterpret_cast<char*>(rec + pos[u]);       // two full-sector stores per record
                bfe_st256(dst, b.c00, b.c10, b.c01, b.c11);
                bfe_st256(dst + 32, c1, s1, aux[u], __longlong_as_double((long long)cp));
                inv[idx[u]] = pos[u];             // original index -> sorted slot (coalesced)
                r_orig[idx[u]] = r;
            }
        }
    }
    BFE_TRACE_PT(1, 1);
}

// ---------------------------------------------------------------------------
// STABLE counting sort without global atomics (option "sort_stable", default 1; north_star (d): "no global atomics on
// the hot loop").  The slot claims of eof_cell_scatter_kernel above are integer atomics on global counters: correct, but
// the order of the records inside a cell -- and with it the FP64 summation order of the deposit -- follows the atomics'
// arrival order, so the coefficients were reproducible to ~1e-16 only.  Here the position of every particle is a pure
// function of the input:
//   eof_tile_hist_kernel    : the set is cut into TILES of 1024 U consecutive particles (U <= 8, about one tile per SM);
//                             one CTA per tile builds the tile's cell histogram in shared memory (shared-memory integer
//                             counters: a count does not depend on the order of its increments) and writes it as row
//                             H[tile][cell] (coalesced);
//   eof_tile_colscan_kernel : thread per cell: exclusive prefix over the tiles in place, H[tile][cell] = number of
//                             particles of that cell in EARLIER tiles; the last CTA scans the cell totals -> cell_start;
//   eof_tile_scatter_kernel : one CTA per tile: shared-memory counting sort of the tile's particle ids by cell (the
//                             provisional order inside a cell comes from shared-memory atomics), then every particle counts
//                             the members of its own (tile, cell) group with a SMALLER id: its rank, independent of the
//                             provisional order.  position = cell_start[cell] + H[tile][cell] + rank: the sort is stable
//                             (ascending particle index inside every cell) and bit-reproducible, and so is everything
//                             downstream of it.  Then the 64-byte record is built and stored as before.
// (512-thread tile CTAs, two per SM, so that tiles of two streams share an SM: measured slower -- scatter 45.7 vs 37.7 us
// per 10^6, step 0.149 vs 0.144 ms -- the per-tile fixed work, clearing and scanning 8192 counters, doubles.)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
eof_tile_hist_kernel(EofGeom g, int ncell, int64_t n, int tile, const double* __restrict__ x, const double* __restrict__ y,
                     const double* __restrict__ z, int* __restrict__ H, int* __restrict__ cellid) {
    extern __shared__ int s_hist[];
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) s_hist[c] = 0;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * tile;
    // two particles per pass: six independent loads in flight per thread (four per pass, fully unrolled, was slower)
    constexpr int PB = 2;
    for (int o0 = threadIdx.x; o0 < tile; o0 += PB * 1024) {
        double px[PB], py[PB], pz[PB];
        bool on[PB];
#pragma unroll
        for (int v = 0; v < PB; ++v) {
            const int o = o0 + v * 1024;
            const int64_t i = base + o;
            on[v] = (o < tile) && (i < n);
            px[v] = on[v] ? __ldg(x + i) : 1.0; py[v] = on[v] ? __ldg(y + i) : 0.0; pz[v] = on[v] ? __ldg(z + i) : 0.0;
        }
#pragma unroll
        for (int v = 0; v < PB; ++v) {
            if (!on[v]) continue;
            int cell;
            if (!bfe_eof_cell_fast(g, px[v], py[v], pz[v], cell)) {   // near a cell edge / unusual input: exact FP64 index
                const double r = sqrt(px[v] * px[v] + py[v] * py[v] + 1.e-10);
                cell = bfe_eof_bin(g, r, pz[v]).cell;
            }
            atomicAdd(&s_hist[cell], 1);                              // shared-memory integer count
            cellid[base + o0 + v * 1024] = cell;
        }
    }
    __syncthreads();
    int* row = H + (size_t)blockIdx.x * ncell;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) row[c] = s_hist[c];
}

// 64 cells x 16 tile groups per CTA (128 CTAs for the 128 x 64 table): a thread loads its share of a cell's column -- up to
// 16 rows, all in flight together, kept in registers -- the 16 partial sums of a cell are exchanged through shared memory,
// and the thread writes its rows back as exclusive prefixes.  (First version: 256 cells x 4 groups = 32 CTAs, two passes
// over the column with 35 dependent-latency loads each: 16 us against 2 us of L2 traffic.)
#define BFE_COLSCAN_CW 64
#define BFE_COLSCAN_TG 16
__global__ void __launch_bounds__(1024)
eof_tile_colscan_kernel(int ncell, int ntile, int* __restrict__ H, int* __restrict__ total, int* __restrict__ cell_start,
                        unsigned int* __restrict__ counter) {
    constexpr int CW = BFE_COLSCAN_CW, TG = BFE_COLSCAN_TG, KEEP = 16;
    __shared__ int s_h[1024];
    __shared__ int s_part[TG][CW];
    __shared__ int s_excl[CW];
    __shared__ int s_wsum[32];
    __shared__ bool s_last;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    const int cl = threadIdx.x & (CW - 1), q = threadIdx.x / CW;
    const int c = blockIdx.x * CW + cl;
    const int tq = (ntile + TG - 1) / TG, t0 = q * tq, t1 = min(ntile, t0 + tq);
    const bool inreg = tq <= KEEP;                              // warp-uniform
    int v[KEEP];
    int sum = 0;
    if (c < ncell) {
        if (inreg) {
#pragma unroll
            for (int u = 0; u < KEEP; ++u) v[u] = (t0 + u < t1) ? __ldcg(H + (size_t)(t0 + u) * ncell + c) : 0;
#pragma unroll
            for (int u = 0; u < KEEP; ++u) sum += v[u];
        } else {
            int t = t0;
            for (; t + 7 < t1; t += 8) {                          // eight independent loads in flight
                int w[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) w[u] = __ldcg(H + (size_t)(t + u) * ncell + c);
#pragma unroll
                for (int u = 0; u < 8; ++u) sum += w[u];
            }
            for (; t < t1; ++t) sum += __ldcg(H + (size_t)t * ncell + c);
        }
    }
    s_part[q][cl] = sum;
    __syncthreads();
    if (c < ncell) {
        int run = 0;
        for (int k = 0; k < q; ++k) run += s_part[k][cl];
        if (inreg) {
#pragma unroll
            for (int u = 0; u < KEEP; ++u)
                if (t0 + u < t1) { H[(size_t)(t0 + u) * ncell + c] = run; run += v[u]; }
        } else {
            int t = t0;
            for (; t + 7 < t1; t += 8) {
                int w[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) w[u] = __ldcg(H + (size_t)(t + u) * ncell + c);
#pragma unroll
                for (int u = 0; u < 8; ++u) { H[(size_t)(t + u) * ncell + c] = run; run += w[u]; }
            }
            for (; t < t1; ++t) { const int w = __ldcg(H + (size_t)t * ncell + c); H[(size_t)t * ncell + c] = run; run += w; }
        }
    }
    // exclusive prefix of this CTA's CW cell totals (two warps), block sum published
    if (threadIdx.x < CW) {
        int tot = 0;
        if (c < ncell)
            for (int k = 0; k < TG; ++k) tot += s_part[k][threadIdx.x];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        int incl = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int w = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += w; }
        if (lane == 31) s_wsum[warp] = incl;
        s_excl[threadIdx.x] = incl - tot;                   // exclusive inside the warp
        s_h[threadIdx.x] = tot;
    }
    __syncthreads();
    if (threadIdx.x < CW) {
        const int warp = threadIdx.x >> 5;
        int add = 0;
        for (int k = 0; k < warp; ++k) add += s_wsum[k];
        if (c < ncell) cell_start[c] = s_excl[threadIdx.x] + add;          // local exclusive prefix; block prefix added below
        if (threadIdx.x == CW - 1) total[blockIdx.x] = s_excl[CW - 1] + add + s_h[CW - 1];   // block sum (total[] re-used: nblk <= ncell)
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);       // one integer add per CTA: not on the particle loop
    __syncthreads();
    if (s_last) {
        __threadfence();
        // block prefixes (<= 1024 blocks) by one warp-shuffle scan, then one coalesced pass that adds them
        const int nb = gridDim.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int bv = (int)threadIdx.x < nb ? __ldcg(total + threadIdx.x) : 0;
        int incl = bv;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int w = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += w; }
        __syncthreads();
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_wsum[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += u; }
            s_wsum[lane] = w;
        }
        __syncthreads();
        s_h[threadIdx.x] = incl - bv + (warp > 0 ? s_wsum[warp - 1] : 0);      // exclusive prefix of block threadIdx.x
        __syncthreads();
        for (int c0 = threadIdx.x; c0 < ncell; c0 += 8 * 1024) {       // eight loads in flight, then the stores (same array)
            int w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int cc = c0 + u * 1024; w[u] = (cc < ncell) ? __ldcg(cell_start + cc) : 0; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int cc = c0 + u * 1024; if (cc < ncell) cell_start[cc] = w[u] + s_h[cc / CW]; }
        }
        if ((int)threadIdx.x == nb - 1) cell_start[ncell] = s_h[nb - 1] + bv;   // particle count
        if (threadIdx.x == 0) *counter = 0u;
        if ((int)threadIdx.x < nb) total[threadIdx.x] = 0;                       // hist / total array left clear, as the other paths expect
    }
}

// the column scan for another table of bins (the SL radial-bin sort, bfe_sl_sort.cu); nbin / 64 <= 1024 blocks
int bfe_tile_colscan(int nbin, int ntile, int* H, int* total, int* bin_start, unsigned int* counter, cudaStream_t stream) {
    eof_tile_colscan_kernel<<<(nbin + BFE_COLSCAN_CW - 1) / BFE_COLSCAN_CW, 1024, 0, stream>>>(nbin, ntile, H, total, bin_start, counter);
    BFE_LAUNCH_CHECK("eof_tile_colscan_kernel");
    return BFE_OK;
}

__global__ void __launch_bounds__(1024)
eof_tile_scatter_kernel(EofGeom g, int ncell, int64_t n, int tile, const double* __restrict__ x, const double* __restrict__ y,
                        const double* __restrict__ z, const double* __restrict__ mass, const int* __restrict__ cell_start,
                        const int* __restrict__ H, EofRec* __restrict__ rec, int* __restrict__ inv, double* __restrict__ r_orig) {
    extern __shared__ int s_mem[];
    __shared__ int s_wsum[32];
    // ONE cell-sized array: counts -> exclusive starts (in place) -> group ENDS (the placement advances each start to the
    // end of its group, which is the start of the next one), so group c is [c ? s_cnt[c-1] : 0, s_cnt[c])
    int* s_cnt = s_mem;                                   // [per * 1024]
    const int per = (ncell + 1023) / 1024;
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_mem + per * 1024);   // [tile] ids ordered by cell
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < per * 1024; c += 1024) s_cnt[c] = 0;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * tile;
    const int nu = tile >> 10;                            // particles per thread (<= 8)
    int cell[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        cell[u] = -1;
        if (u < nu) {
            const int64_t i = base + u * 1024 + tid;
            if (i < n) { cell[u] = inv[i]; atomicAdd(&s_cnt[cell[u]], 1); }      // cell id left here by the histogram kernel
        }
    }
    __syncthreads();
    // exclusive scan of the counts over the cells, in place (thread = `per` consecutive cells, two levels of warp shuffles)
    {
        const int lo = tid * per;
        int sum = 0;
        for (int k = 0; k < per; ++k) sum += s_cnt[lo + k];
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_wsum[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += v; }
            s_wsum[lane] = w;
        }
        __syncthreads();
        int run = incl - sum + (warp > 0 ? s_wsum[warp - 1] : 0);
        for (int k = 0; k < per; ++k) { const int h = s_cnt[lo + k]; s_cnt[lo + k] = run; run += h; }
    }
    __syncthreads();
    // provisional placement of the ids (order inside a cell: arrival order of the shared-memory atomics)
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (cell[u] >= 0) s_list[atomicAdd(&s_cnt[cell[u]], 1)] = (unsigned short)(u * 1024 + tid);
    __syncthreads();
    // rank = members of the own (tile, cell) group with a smaller id; record; stores.  One particle per pass, the loads of
    // the NEXT pass (coordinates, mass, cell_start, tile offset) issued before the arithmetic of this one: with two
    // particles per pass and no look-ahead every pass exposed one DRAM latency (ncu: issue slots 46 % busy, 1 CTA of 32
    // warps per SM).  The loop stays rolled: unrolled eight times it was 9800 instructions and slower (instruction cache).
    struct Pre { double x, y, z, m; int g0, g1; };
    auto cell_at = [&](int u) {                           // cell[] stays in registers: constant indices only
        int c = cell[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) c = (u == k) ? cell[k] : c;
        return c;
    };
    auto fetch = [&](int u, int c) {
        Pre q;
        const bool on = c >= 0;
        const int64_t i = base + u * 1024 + tid;
        q.x = on ? __ldg(x + i) : 1.0; q.y = on ? __ldg(y + i) : 0.0; q.z = on ? __ldg(z + i) : 0.0;
        q.m = (on && mass) ? __ldg(mass + i) : 0.0;
        q.g0 = on ? __ldg(cell_start + c) : 0;
        q.g1 = on ? __ldg(H + (size_t)blockIdx.x * ncell + c) : 0;
        return q;
    };
    Pre nxt = fetch(0, cell[0]);
#pragma unroll 1
    for (int u = 0; u < nu; ++u) {
        const Pre cur = nxt;
        const int c = cell_at(u);
        if (u + 1 < nu) nxt = fetch(u + 1, cell_at(u + 1));
        if (c < 0) continue;
        const int64_t idx = base + u * 1024 + tid;
        const int lo = c ? s_cnt[c - 1] : 0, hi = s_cnt[c];
        const int me = u * 1024 + tid;
        // ids are < 2^13 in 16-bit slots: four per 64-bit shared-memory load, compared without a SIMD instruction (sm_100
        // emulates __vcmpltu2 with ~6 logic ops): for x, me < 2^15 bit 15 of (me + 0x7fff - x) is set iff x < me, and
        // neither half borrows from the other.
        int rank = 0;
        int k = lo;
        for (; (k & 3) && k < hi; ++k) rank += (s_list[k] < me) ? 1 : 0;
        const unsigned int K = (unsigned int)me * 0x00010001u + 0x7fff7fffu;
        unsigned int acc = 0u;
        for (; k + 3 < hi; k += 4) {
            const uint2 w = *reinterpret_cast<const uint2*>(s_list + k);
            acc += (((K - w.x) >> 15) & 0x00010001u) + (((K - w.y) >> 15) & 0x00010001u);
        }
        rank += (int)((acc & 0xffffu) + (acc >> 16));
        for (; k < hi; ++k) rank += (s_list[k] < me) ? 1 : 0;
        const int pos = cur.g0 + cur.g1 + rank;
        const double r = sqrt(cur.x * cur.x + cur.y * cur.y + 1.e-10);   // eof.py:531 / 1070
        const EofBin b = bfe_eof_bin(g, r, cur.z);                       // b.cell == c
        double c1, s1;
        bfe_cossin_phi(cur.x, cur.y, c1, s1);
        const unsigned long long cp = ((unsigned long long)(unsigned int)c << 32) | (unsigned long long)(unsigned int)idx;
        char* dst = reinterpret_cast<char*>(rec + pos);                  // two full-sector stores per record
        bfe_st256(dst, b.c00, b.c10, b.c01, b.c11);
        bfe_st256(dst + 32, c1, s1, cur.m, __longlong_as_double((long long)cp));
        inv[idx] = pos;                                                  // original index -> sorted slot (coalesced)
        r_orig[idx] = r;
    }
}

// ---------------------------------------------------------------------------
// deposit + contract over sorted records, in two kernels.
//
// eof_segsum_kernel -- a WARP owns a task of 128 consecutive sorted records (dynamic task queue) and walks it
// 32 records at a time:
//   1. lanes = records: the 64-B record (prefetched one batch ahead) is expanded in the warp's shared-memory
//      slab to 8 A rows (the 4 mass-weighted corner weights w_k, and w_k cos 4phi) and 7 B rows
//      (cos/sin phi, 2phi, 3phi and sin 4phi; the eighth column is the constant 1);
//   2. the per-run sums are a (8 x L)(L x 8) product done on the FP64 tensor cores, ONE mma.sync.m8n8k4 per
//      4 records: D[k][.] = sum_p w_k {1, c1, s1, c2, s2, c3, s3, s4} and D[4+k][.] = the same with w_k c4,
//      from which the harmonics 4..6 follow by the product formulas (cos 4x cos x = (cos 5x + cos 3x)/2, ...):
//      all 64 outputs of the tile carry information, 52 sums come out of it.  Records outside the current run
//      are masked to zero in the A fragment, so a run boundary costs at most one extra k-step;
//   3. when the cell id changes (warp-uniform) the tile is stored (512 B, coalesced) to the segment slot
//      cell + task -- unique and monotone along the sorted array -- and the accumulators are cleared.
// eof_node_contract_kernel -- thread per (m,n,trig) channel j: for every non-empty cell the segment tiles of
// the cell are summed, the 52 sums S[k][ch] derived, and  acc_j += sum_k T[node_k][j] S[k][ch(j)]  with the four
// node rows read coalesced; per-CTA partials, last-CTA reduce (4 row groups x 16 rows in flight).
// The first version of this pass (one kernel: runs flushed into 234 register accumulators by the summing
// warp, table rows by TMA) was bound by the flushes of short runs: slowest warp 90 k cycles against a median
// of 48 k, plus a 14 us serial epilogue (per-warp cycle counters, profiles/r01_summary.md section 11).
// No floating-point atomics anywhere.
// ---------------------------------------------------------------------------
struct SegSum {
    static constexpr int TASK = 128;              // records per warp task
    static constexpr int NW = 8;                  // warps per CTA
    static constexpr int RS = 36;                 // slab row stride (doubles): conflict-free fragment loads
    static constexpr int NROW = 15;               // 8 A rows + 7 B rows
};

__global__ void __launch_bounds__(256, 4)
eof_segsum_kernel(int64_t n, const EofRec* __restrict__ rec, double* __restrict__ seg,
                  unsigned int* __restrict__ counter) {
    constexpr int TASK = SegSum::TASK, NW = SegSum::NW, RS = SegSum::RS, SLAB = SegSum::NROW * SegSum::RS;
    __shared__ double s_slab[NW * SLAB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fg = lane >> 2, fj = lane & 3;      // mma fragment coordinates: group, thread-in-group
    double* val = s_slab + warp * SLAB;           // val[row * RS + record]
    const double* arow = val + fg * RS;           // A[row fg][.]
    const double* brow = val + (7 + fg) * RS;     // B[.][col fg] (col 0 is the constant 1)
    unsigned int* task_counter = counter + 1;
    const int64_t ntasks = (n + TASK - 1) / TASK;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    BFE_TRACE_PT(2, 0);
    for (;;) {
        int64_t task = 0;
        if (lane == 0) task = (int64_t)atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntasks) break;
        task = ntasks - 1 - task;                 // last task first: the short runs of the sparse outskirts are at the end
        const int64_t t0 = task * TASK;
        const int tcnt = (int)((n - t0) < TASK ? (n - t0) : TASK);
        double d0 = 0.0, d1 = 0.0;                // D[fg][2 fj], D[fg][2 fj + 1]
        int cur = -1;                             // cell of the run being summed (-1: none)
        double2 ra, rb, rc, rd;
        {
            const bool on = lane < ((tcnt < 32) ? tcnt : 32);
            const char* src = reinterpret_cast<const char*>(rec + t0 + (on ? lane : 0));
            bfe_ld256_nc(src, ra.x, ra.y, rb.x, rb.y); bfe_ld256_nc(src + 32, rc.x, rc.y, rd.x, rd.y);
        }
        for (int b0 = 0; b0 < tcnt; b0 += 32) {
            const int bcnt = (tcnt - b0) < 32 ? (tcnt - b0) : 32;
            __syncwarp();
            int mycell = -3;
            {
                // lanes >= bcnt park zero weights so that k-steps may read the whole 32-record slab
                const bool on = lane < bcnt;
                const double m = on ? rd.x : 0.0;
                const double c1 = rc.x, s1 = rc.y;
                const double c2 = c1 * c1 - s1 * s1, s2 = 2.0 * c1 * s1;
                const double c3 = c2 * c1 - s2 * s1, s3 = s2 * c1 + c2 * s1;
                const double c4 = c2 * c2 - s2 * s2, s4 = 2.0 * c2 * s2;
                const double w0 = ra.x * m, w1 = ra.y * m, w2 = rb.x * m, w3 = rb.y * m;
                val[0 * RS + lane] = w0; val[1 * RS + lane] = w1; val[2 * RS + lane] = w2; val[3 * RS + lane] = w3;
                val[4 * RS + lane] = w0 * c4; val[5 * RS + lane] = w1 * c4;
                val[6 * RS + lane] = w2 * c4; val[7 * RS + lane] = w3 * c4;
                val[8 * RS + lane] = c1; val[9 * RS + lane] = s1; val[10 * RS + lane] = c2; val[11 * RS + lane] = s2;
                val[12 * RS + lane] = c3; val[13 * RS + lane] = s3; val[14 * RS + lane] = s4;
                if (on) mycell = (int)((unsigned long long)__double_as_longlong(rd.y) >> 32);
            }
            if (b0 + 32 < tcnt) {
                const bool on = (b0 + 32 + lane) < tcnt;
                const char* src = reinterpret_cast<const char*>(rec + t0 + b0 + 32 + (on ? lane : 0));
                bfe_ld256_nc(src, ra.x, ra.y, rb.x, rb.y); bfe_ld256_nc(src + 32, rc.x, rc.y, rd.x, rd.y);
            }
            // run boundaries inside this batch (bit p set: record p starts a new run)
            int prevcell = __shfl_up_sync(0xffffffffu, mycell, 1);
            if (lane == 0) prevcell = cur;
            const unsigned int bmask = __ballot_sync(0xffffffffu, (lane < bcnt) && (mycell != prevcell));
            __syncwarp();
            int p = 0;
            while (p < bcnt) {
                if ((bmask >> p) & 1u) {
                    if (cur >= 0) {
                        reinterpret_cast<double2*>(seg + ((size_t)cur + (size_t)task) * 64)[lane] = make_double2(d0, d1);
                        d0 = 0.0; d1 = 0.0;
                    }
                    cur = __shfl_sync(0xffffffffu, mycell, p);
                }
                // end of this run within the batch: next boundary after p, or bcnt
                const unsigned int later = (p < 31) ? (bmask >> (p + 1)) : 0u;
                const int qend = later ? (p + __ffs(later)) : bcnt;
                // k-steps of 4 records covering [p, qend); records outside the run are masked in A
                for (int ks = p >> 2; ks * 4 < qend; ++ks) {
                    const int r = ks * 4 + fj;
                    const bool in = (r >= p) && (r < qend);
                    const double a = in ? arow[r] : 0.0;
                    const double b = fg ? brow[r] : 1.0;
                    bfe_dmma_m8n8k4(d0, d1, a, b);
                }
                p = qend;
            }
        }
        if (cur >= 0)
            reinterpret_cast<double2*>(seg + ((size_t)cur + (size_t)task) * 64)[lane] = make_double2(d0, d1);
    }
    // the last CTA to finish re-arms the task queue
    BFE_TRACE_PT(2, 1);
    __syncthreads();
    BFE_TRACE_PT(2, 2);
    if (tid == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) { counter[0] = 0u; counter[1] = 0u; }
    }
    BFE_TRACE_PT(2, 3);
}

// A sum S[k][ch] from a raw 8x8 tile R (row-major; rows 0-3 = sum_p w_k {1, c1, s1, c2, s2, c3, s3, s4}, rows
// 4-7 the same with w_k c4):  S = c_lo R[k][i_lo] + c_hi R[4+k][i_hi], the selectors depending only on the harmonic.
struct SegSel { int i_lo, i_hi; double c_lo, c_hi; };
__device__ __forceinline__ SegSel bfe_seg_selector(int m, bool is_sin) {
    SegSel s; s.i_lo = 0; s.i_hi = 0; s.c_lo = 1.0; s.c_hi = 0.0;
    if (!is_sin) {
        if (m <= 3) s.i_lo = (m == 0) ? 0 : (2 * m - 1);                     // 1, c1, c2, c3
        else if (m == 4) { s.c_lo = 0.0; s.c_hi = 1.0; s.i_hi = 0; }          // sum w c4
        else { s.c_hi = 2.0; s.i_hi = 2 * (m - 4) - 1; s.c_lo = -1.0; s.i_lo = 2 * (8 - m) - 1; }   // c5 = 2 c4 c1 - c3, c6 = 2 c4 c2 - c2
    } else {
        if (m <= 3) s.i_lo = 2 * m;                                           // s1, s2, s3
        else if (m == 4) s.i_lo = 7;                                          // s4
        else { s.c_hi = 2.0; s.i_hi = 2 * (m - 4); s.c_lo = 1.0; s.i_lo = 2 * (8 - m); }            // s5 = 2 c4 s1 + s3, s6 = 2 c4 s2 + s2
    }
    return s;
}

__global__ void __launch_bounds__(256)
eof_node_contract_kernel(EofGeom g, const double* __restrict__ t_acc, int nch, int nch_pad, int ncell,
                         const int* __restrict__ cell_start, const double* __restrict__ seg,
                         double* __restrict__ partial, unsigned int* __restrict__ counter,
                         double* __restrict__ cos_out, double* __restrict__ sin_out) {
    constexpr int CB = 8;                         // cells staged per pass
    constexpr int TASK = SegSum::TASK;
    constexpr int GROUP = 16;                     // CTAs per first-level reduce group
    __shared__ double s_raw[CB][64];
    __shared__ int s_flag[2];
    const int tid = threadIdx.x;
    const int ncos = (g.mmax + 1) * g.norder;
    // this thread's channel j = tid: harmonic and the selectors of its sums
    SegSel sel = bfe_seg_selector(0, false);
    if (tid < nch) {
        const bool is_sin = tid >= ncos;
        const int m = is_sin ? (1 + (tid - ncos) / g.norder) : (tid / g.norder);
        sel = bfe_seg_selector(m, is_sin);
    }
    const int rowstep = nch_pad, colstep = g.ny1 * nch_pad;
    const double* tcol = t_acc + (tid < nch ? tid : 0);
    double acc = 0.0;
    const int nblk = (ncell + CB - 1) / CB;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    BFE_TRACE_PT(3, 0);
    for (int blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        // The CB cells of a block are nblk apart (same iy, ix 16 apart for the 128 x 64 table): consecutive cells
        // are all hot or all empty, and with 8-cell blocks handed out with a stride that is a multiple of 8 a
        // quarter of the CTAs got every hot block (device timeline: slowest CTA 15 us, median 6 us).
        // Every thread reads the run offsets itself (broadcast loads), so the table rows of the non-empty cells
        // can be requested at once, together with the segment tiles.
        int cs0[CB], cs1[CB];
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            const int cell = c * nblk + blk;
            cs0[c] = (cell < ncell) ? __ldg(cell_start + cell) : 0;
            cs1[c] = (cell < ncell) ? __ldg(cell_start + cell + 1) : 0;
        }
        double t00[CB], t10[CB], t01[CB], t11[CB];
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            if (cs1[c] > cs0[c]) {
                const int cell = c * nblk + blk;
                const int ix = cell / g.numy, iy = cell - ix * g.numy;
                const double* base = tcol + (size_t)(ix * g.ny1 + iy) * nch_pad;
                t00[c] = __ldg(base); t10[c] = __ldg(base + colstep);
                t01[c] = __ldg(base + rowstep); t11[c] = __ldg(base + colstep + rowstep);
            } else { t00[c] = 0.0; t10[c] = 0.0; t01[c] = 0.0; t11[c] = 0.0; }
        }
        // ---- sum the segment tiles of each cell of this block (thread = one tile entry of one cell)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int c = (tid >> 6) + 4 * h, ent = tid & 63;
            const int cell = c * nblk + blk;
            int s = 0, e = 0;
#pragma unroll
            for (int q = 0; q < CB; ++q) if (q == c) { s = cs0[q]; e = cs1[q]; }
            double v = 0.0;
            if (e > s) {
                const int ta = s / TASK, tb = (e - 1) / TASK;
                const double* sp = seg + ((size_t)cell + (size_t)ta) * 64 + ent;
                for (int t = ta; t <= tb; t += 8, sp += 8 * 64) {               // 8 independent loads in flight
                    double u[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) u[q] = (t + q <= tb) ? __ldcg(sp + q * 64) : 0.0;
                    v += ((u[0] + u[1]) + (u[2] + u[3])) + ((u[4] + u[5]) + (u[6] + u[7]));
                }
            }
            s_raw[c][ent] = v;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CB; ++c) {
            if (cs1[c] > cs0[c]) {
                const double* lo = &s_raw[c][sel.i_lo];
                const double* hi = &s_raw[c][32 + sel.i_hi];
                const double S0 = sel.c_lo * lo[0] + sel.c_hi * hi[0], S1 = sel.c_lo * lo[8] + sel.c_hi * hi[8];
                const double S2 = sel.c_lo * lo[16] + sel.c_hi * hi[16], S3 = sel.c_lo * lo[24] + sel.c_hi * hi[24];
                acc += t00[c] * S0 + t10[c] * S1 + t01[c] * S2 + t11[c] * S3;
            }
        }
        __syncthreads();
    }
    // ---- two-level reduce of the per-CTA partials in fixed order (deterministic): the last CTA of each group of
    // 16 sums its group's rows, the last group to finish sums the group rows.  A single last CTA reading all 296
    // rows through one SM cost ~12 us of serial tail (ncu: sm__cycles_active max vs avg).
    BFE_TRACE_PT(3, 1);
    const int ngroups = ((int)gridDim.x + GROUP - 1) / GROUP;
    const int grp = blockIdx.x / GROUP;
    const int gfirst = grp * GROUP;
    const int gsize = ((int)gridDim.x - gfirst) < GROUP ? ((int)gridDim.x - gfirst) : GROUP;
    double* grow = partial + (size_t)gridDim.x * nch_pad;                 // [ngroups][nch_pad]
    if (tid < nch_pad) partial[(size_t)blockIdx.x * nch_pad + tid] = (tid < nch) ? acc : 0.0;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_flag[0] = (atomicAdd(counter + 64 + grp, 1u) == (unsigned int)(gsize - 1));
    __syncthreads();
    BFE_TRACE_PT(3, 2);
    if (!s_flag[0]) return;
    __threadfence();
    if (tid < nch_pad) {
        double v[GROUP];
#pragma unroll
        for (int r = 0; r < GROUP; ++r) v[r] = (r < gsize) ? __ldcg(partial + (size_t)(gfirst + r) * nch_pad + tid) : 0.0;
        double sum = 0.0;
#pragma unroll
        for (int r = 0; r < GROUP; ++r) sum += v[r];
        grow[(size_t)grp * nch_pad + tid] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        counter[64 + grp] = 0u;
        s_flag[1] = (atomicAdd(counter, 1u) == (unsigned int)(ngroups - 1));
    }
    __syncthreads();
    BFE_TRACE_PT(3, 3);
    if (!s_flag[1]) return;
    __threadfence();
    if (tid < nch) {
        double sum = 0.0;
        for (int r0 = 0; r0 < ngroups; r0 += 16) {
            double v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = (r0 + r < ngroups) ? __ldcg(grow + (size_t)(r0 + r) * nch_pad + tid) : 0.0;
#pragma unroll
            for (int r = 0; r < 16; ++r) sum += v[r];
        }
        sum *= BFE_FOURPI_NEG;
        if (tid < ncos) cos_out[tid] = sum;
        else            sin_out[tid - ncos + g.norder] = sum;
    }
    for (int k = tid; k < g.norder; k += 256) sin_out[k] = 0.0;
    if (tid == 0) counter[0] = 0u;
    BFE_TRACE_PT(3, 4);
}

// ---------------------------------------------------------------------------
// force evaluation in sorted order: thread per sorted record, a warp's lanes share the contracted-grid
// rows (warp-uniform addresses).  Results go to a 48-byte AoS slot in SORTED order (coalesced); a second
// kernel in the caller's particle order gathers them through the inverse permutation and writes the six
// SoA outputs coalesced.  (Scattering six 8-byte stores per thread straight from this kernel cost more
// than the evaluation itself: partial-sector writes and their fills, ncu profiles/.)
// ---------------------------------------------------------------------------
template <int MCAP>
__global__ void __launch_bounds__(128)
eof_force_sorted_kernel(EofGeom g, const double* __restrict__ G, int gstride, int64_t n,
                        const EofRec* __restrict__ rec, double2* __restrict__ tmp) {
    // each CTA takes one contiguous slice of the sorted records: an SM then sees a contiguous range
    // of cells and the contracted-grid rows stay in its L1
    const int64_t per = ((n + gridDim.x - 1) / gridDim.x + 127) / 128 * 128;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    const int64_t lo = (int64_t)blockIdx.x * per;
    const int64_t hi = (lo + per) < n ? (lo + per) : n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const double2* src = reinterpret_cast<const double2*>(rec + i);
        double2 a = __ldg(src), bb = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
        unsigned long long cp = (unsigned long long)__double_as_longlong(d.y);
        const int cell = (int)(cp >> 32);
        EofBin b;
        const int ix = cell / g.numy, iy = cell - ix * g.numy;
        b.node = ix * g.ny1 + iy;
        b.c00 = a.x; b.c10 = a.y; b.c01 = bb.x; b.c11 = bb.y;
        EofField f = bfe_eof_eval<MCAP>(g, G, gstride, b, c.x, c.y);
        double2* dst = tmp + 3 * i;
        dst[0] = make_double2(f.p0, f.p);
        dst[1] = make_double2(f.fr, f.fp);
        dst[2] = make_double2(f.fz, 0.0);
    }
}

// ---------------------------------------------------------------------------
// The same evaluation on the FP64 tensor cores.  For the records of one cell the 42 interpolated grid values
// V[p][c] = sum_k w_k(p) G[node_k][c] are a (particles x 4 corners)(4 x 42) product: mma.sync.m8n8k4 with
// rows = 8 consecutive sorted records, k = the four corners, 6 n-tiles of 8 columns.  The B fragments (6 doubles
// per lane) are loaded once per RUN of equal cells; a group of 8 records that straddles runs is done run by run
// with the A rows of the other runs masked to zero, all into the same accumulators (each row belongs to exactly
// one run).  Per record the per-lane kernel above moves 1344 B of grid rows into registers (84 LDG.128, L1
// data-return bound, ncu profiles/); here it is 32 B of weights.
// Column layout: tile t holds field t/2 (potential, radial, vertical) and harmonics m = 4 (t%2) + col/2 as
// (cos, sin) pairs, so a lane's two accumulators of a tile are one (cos, sin) pair of ONE harmonic and every
// lane needs the trig factors of just two harmonics (m = j and m = 4 + j, j = lane % 4).  The five outputs of
// a record are then spread over the 4 lanes of its row and combined with two butterfly shuffles.
// ---------------------------------------------------------------------------
// The records are streamed through shared memory by TMA bulk copies (cp.async.bulk + mbarrier), 32 records
// (2 kB) per copy, double-buffered per warp: with register prefetch one group ahead a warp had only 512 B in
// flight (latency-bound, ncu profiles/).  Work is handed out dynamically in tickets of 128 records, LAST
// ticket first: the sparse outskirts of the sorted array (short runs, one dependent B-fragment load each) are
// at its end, and with a static split the warps that got them finished 50 % after the average.
// Epilogue of one group of 8 records: the lane (row, jj) holds the interpolated (cos, sin) pairs of harmonics jj
// and 4 + jj for the three fields (d[2 f][.] / d[2 f + 1][.]).  Trig factors as powers of z = cos phi + i sin phi,
// the lane's share of the four sums, then a TRANSPOSED 4-lane reduction: after the xor-1 step even lanes hold
// (p, fz) and odd lanes (fr, fp), after the xor-2 step lane 0 holds p, lane 1 fr, lane 2 fz, lane 3 fp -- three
// 64-bit shuffles instead of the eight of a butterfly on all four values, and every lane stores its own piece.
struct ForceLaneConst { double mlo, mhi; bool hi_on, bit0, bit1; int jj; };
__device__ __forceinline__ void bfe_force_group_epilogue(const double (&d)[6][2], double2 cs1, const ForceLaneConst& k,
                                                         bool on, double* __restrict__ slot) {
    const double c2 = cs1.x * cs1.x - cs1.y * cs1.y, s2 = 2.0 * cs1.x * cs1.y;
    const double c4 = c2 * c2 - s2 * s2, s4 = 2.0 * c2 * s2;
    const double ac = k.bit0 ? cs1.x : 1.0, as = k.bit0 ? cs1.y : 0.0;
    const double clo = k.bit1 ? (ac * c2 - as * s2) : ac, slo = k.bit1 ? (as * c2 + ac * s2) : as;
    const double chi = c4 * clo - s4 * slo, shi = s4 * clo + c4 * slo;
    // tiles 0,1: potential pairs; 2,3: radial force; 4,5: vertical force
    double pp = (k.jj == 0) ? 0.0 : (clo * d[0][0] + slo * d[0][1]);        // m = 0 goes to p0, not p
    double fp = k.mlo * (slo * d[0][0] - clo * d[0][1]);
    double fr = clo * d[2][0] + slo * d[2][1];
    double fz = clo * d[4][0] + slo * d[4][1];
    if (k.hi_on) {
        pp += chi * d[1][0] + shi * d[1][1];
        fp += k.mhi * (shi * d[1][0] - chi * d[1][1]);
        fr += chi * d[3][0] + shi * d[3][1];
        fz += chi * d[5][0] + shi * d[5][1];
    }
    const double ra = __shfl_xor_sync(0xffffffffu, k.bit0 ? pp : fr, 1);
    const double rb = __shfl_xor_sync(0xffffffffu, k.bit0 ? fz : fp, 1);
    const double ka = (k.bit0 ? fr : pp) + ra;          // even lanes: p, odd lanes: fr   (sums over a lane pair)
    const double kb = (k.bit0 ? fp : fz) + rb;          // even lanes: fz, odd lanes: fp
    const double rc = __shfl_xor_sync(0xffffffffu, k.bit1 ? ka : kb, 2);
    const double fin = (k.bit1 ? kb : ka) + rc;         // lane 0: p, 1: fr, 2: fz, 3: fp
    if (on) {
        // slot = {p0, p, fr, fp, fz, 0}: all 48 bytes are written (a sector left partly unwritten costs a DRAM
        // read-modify-write when it is evicted)
        if (!k.bit0) *reinterpret_cast<double2*>(slot + 2 * k.jj) = k.bit1 ? make_double2(fin, 0.0) : make_double2(d[0][0], fin);
        else slot[k.bit1 ? 3 : 2] = fin;
    }
}

#ifndef BFE_FMMA_CTAS
#define BFE_FMMA_CTAS 5          // resident CTAs per SM the kernel is compiled for (register cap 65536 / (128 CTAS))
#endif
#ifndef BFE_FMMA_UNROLL
#define BFE_FMMA_UNROLL 2        // groups of the uniform-chunk path unrolled together
#endif
template <int MCAP>
__global__ void __launch_bounds__(128, BFE_FMMA_CTAS)
eof_force_sorted_mma_kernel(EofGeom g, const double* __restrict__ G, int gstride, int64_t n,
                            const EofRec* __restrict__ rec, double2* __restrict__ tmp,
                            unsigned int* __restrict__ counter) {
    static_assert(MCAP <= 6, "harmonics 0..6 fit the 8 pair slots of two n-tiles");
    constexpr int NWARP = 4, CHUNK = 32, TICKET = 128;
    constexpr int SPARSE_RUNS = 6;                 // run starts per 32-record chunk above which the chunk is gathered
    __shared__ __align__(128) double s_buf[NWARP][2][CHUNK * 8];
    __shared__ unsigned long long s_bar[NWARP][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row = lane >> 2, jj = lane & 3;
    const unsigned int bar0 = bfe_smem_u32(&s_bar[warp][0]), bar1 = bfe_smem_u32(&s_bar[warp][1]);
    const unsigned int buf0 = bfe_smem_u32(&s_buf[warp][0][0]), buf1 = bfe_smem_u32(&s_buf[warp][1][0]);
    // Guided tickets: records [nsmall_rec, n) go out in tickets of 128 (4 chunks), LAST ticket first (the sparse
    // outskirts of the sorted array -- short runs, slow chunks -- are at its end); the dense head [0, nsmall_rec)
    // goes out last in single chunks, two per warp of the grid, so that the finish line is one chunk (~0.9 us)
    // wide instead of one ticket (device timeline: last CTA 16 us after the median with uniform tickets).
    const int64_t head = (int64_t)gridDim.x * NWARP * 2 * CHUNK;
    const int64_t nsmall_rec = (head < n ? head : n) / TICKET * TICKET;
    const int nsmall = (int)(nsmall_rec / CHUNK);
    const int nbig = (int)((n - nsmall_rec + TICKET - 1) / TICKET);
    const int ntick = nbig + nsmall;
    unsigned int* ticket_counter = counter + 1;
    BFE_TRACE_PT(5, 0);
    // producer state (meaningful in lane 0): next record / chunks left of the current ticket, the two issued chunks
    int64_t pf_next = 0;
    int pf_left = 0;
    int qb0 = 0, qc0 = 0, qb1 = 0, qc1 = 0;        // record base / count of the chunk in buffer 0 / 1 (count 0: none)
    auto issue = [&](int par) {                    // lane 0 only
        if (pf_left == 0) {
            const int t = (int)atomicAdd(ticket_counter, 1u);
            if (t < nbig) {
                pf_next = nsmall_rec + (int64_t)(nbig - 1 - t) * TICKET;
                const int64_t left = n - pf_next;
                pf_left = (int)(((left < TICKET ? left : TICKET) + CHUNK - 1) / CHUNK);
            } else if (t < ntick) {
                pf_next = (int64_t)(nsmall - 1 - (t - nbig)) * CHUNK;
                pf_left = 1;
            }
        }
        int base = 0, cnt = 0;
        if (pf_left > 0) {
            base = (int)pf_next;
            const int64_t left = n - pf_next;
            cnt = (int)(left < CHUNK ? left : CHUNK);
            pf_next += CHUNK;
            --pf_left;
            const unsigned int bytes = (unsigned int)cnt * 64u;
            bfe_mbar_expect_tx(par ? bar1 : bar0, bytes);
            bfe_bulk_g2s(par ? buf1 : buf0, rec + base, bytes, par ? bar1 : bar0);
        }
        if (par) { qb1 = base; qc1 = cnt; } else { qb0 = base; qc0 = cnt; }
    };
    if (lane == 0) {
        bfe_mbar_init(bar0, 1);
        bfe_mbar_init(bar1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    bfe_pdl_wait();
    bfe_pdl_trigger();
    if (lane == 0) {
        issue(0);
        issue(1);
    }
    __syncwarp();
    // B fragment of this lane: B[k = jj][col = row] of tile t  ->  G[node_k][m*6 + field*2 + cs]
    const int koff = ((jj & 1) ? g.ny1 : 0) + ((jj & 2) ? 1 : 0);    // record order c00, c10, c01, c11
    int boff[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) {
        const int m = 4 * (t & 1) + (row >> 1), field = t >> 1, cs = row & 1;
        boff[t] = (m <= g.mmax) ? (m * 6 + field * 2 + cs) : -1;
    }
    ForceLaneConst kc;
    kc.jj = jj; kc.mlo = (double)jj; kc.mhi = (4 + jj <= g.mmax) ? (double)(4 + jj) : 0.0;
    kc.hi_on = (4 + jj) <= g.mmax; kc.bit0 = jj & 1; kc.bit1 = jj & 2;
    double B[6];
#pragma unroll
    for (int t = 0; t < 6; ++t) B[t] = 0.0;
    int curcell = -1;
    auto load_B = [&](int c) {
        curcell = c;
        const int ix = c / g.numy, iy = c - ix * g.numy;
        const double* gp = G + (size_t)(ix * g.ny1 + iy + koff) * gstride;
#pragma unroll
        for (int t = 0; t < 6; ++t) B[t] = (boff[t] >= 0) ? __ldg(gp + boff[t]) : 0.0;
    };
    unsigned int ph0 = 0u, ph1 = 0u;
    for (int par = 0;; par ^= 1) {
        const int cbase = __shfl_sync(0xffffffffu, par ? qb1 : qb0, 0);
        const int ccnt = __shfl_sync(0xffffffffu, par ? qc1 : qc0, 0);
        if (ccnt == 0) break;
        if (par) { bfe_mbar_wait(bar1, ph1); ph1 ^= 1u; } else { bfe_mbar_wait(bar0, ph0); ph0 ^= 1u; }
        const int* rl = reinterpret_cast<const int*>(&s_buf[warp][par][(lane < ccnt ? lane : 0) * 8]);
        const int mycell = rl[15];
        const int prev = __shfl_up_sync(0xffffffffu, mycell, 1);
        const unsigned int starts = __ballot_sync(0xffffffffu, (lane < ccnt) && (lane == 0 || mycell != prev));
        // chunks made of short runs (sparse outskirts): every run would cost one dependent B-fragment load, ~0.5 us
        // each, serialised in this warp.  Such chunks are evaluated lane per record with the gather formulation
        // instead: 84 independent 16-B loads per lane, all in flight together.
        if (__popc(starts) > SPARSE_RUNS) {
            if (lane < ccnt) {
                const double* r8 = &s_buf[warp][par][lane * 8];
                EofBin b;
                const int ix = mycell / g.numy, iy = mycell - ix * g.numy;
                b.node = ix * g.ny1 + iy; b.cell = mycell;
                b.c00 = r8[0]; b.c10 = r8[1]; b.c01 = r8[2]; b.c11 = r8[3];
                const EofField f = bfe_eof_eval<MCAP>(g, G, gstride, b, r8[4], r8[5]);
                double2* o = tmp + 3 * ((int64_t)cbase + lane);
                o[0] = make_double2(f.p0, f.p);
                o[1] = make_double2(f.fr, f.fp);
                o[2] = make_double2(f.fz, 0.0);
            }
            __syncwarp();
            if (lane == 0) issue(par);
            continue;
        }
        const double* sb = &s_buf[warp][par][row * 8];
        double* slot = reinterpret_cast<double*>(tmp + 3 * ((int64_t)cbase + row));
        if (ccnt == CHUNK && (starts >> 1) == 0u) {
            // ---- the common case (92 % of the chunks of a 10^6-particle disc): 32 records of ONE cell.  No
            // per-group run tests, no tail predicates, and the four groups are independent instruction streams
            // for the scheduler (unrolled).
            const int c0 = __shfl_sync(0xffffffffu, mycell, 0);
            if (c0 != curcell) load_B(c0);
            constexpr int kUnroll = BFE_FMMA_UNROLL;
#pragma unroll kUnroll
            for (int gi = 0; gi < CHUNK / 8; ++gi) {
                const double* rp = sb + gi * 64;
                const double w = rp[jj];
                const double2 cs1 = *reinterpret_cast<const double2*>(rp + 4);
                double d[6][2];
#pragma unroll
                for (int t = 0; t < 6; ++t) { d[t][0] = 0.0; d[t][1] = 0.0; bfe_dmma_m8n8k4(d[t][0], d[t][1], w, B[t]); }
                bfe_force_group_epilogue(d, cs1, kc, true, slot + gi * 48);
            }
        } else {
#pragma unroll 1
            for (int gi = 0; gi < CHUNK / 8; ++gi) {
                if (gi * 8 >= ccnt) break;                                // warp-uniform
                const bool on = (gi * 8 + row) < ccnt;
                const double* rp = sb + gi * 64;
                const double w = on ? rp[jj] : 0.0;
                const double2 cs1 = on ? *reinterpret_cast<const double2*>(rp + 4) : make_double2(1.0, 0.0);
                const int cell = on ? reinterpret_cast<const int*>(rp)[15] : -1;      // high word of cellperm
                double d[6][2];
#pragma unroll
                for (int t = 0; t < 6; ++t) { d[t][0] = 0.0; d[t][1] = 0.0; }
                unsigned int todo = __ballot_sync(0xffffffffu, on);
                while (todo) {
                    const int c = __shfl_sync(0xffffffffu, cell, __ffs(todo) - 1);
                    if (c != curcell) load_B(c);
                    const bool mine = cell == c;
                    const double a = mine ? w : 0.0;
#pragma unroll
                    for (int t = 0; t < 6; ++t) bfe_dmma_m8n8k4(d[t][0], d[t][1], a, B[t]);
                    todo &= ~__ballot_sync(0xffffffffu, mine);
                }
                bfe_force_group_epilogue(d, cs1, kc, on, slot + gi * 48);
            }
        }
        __syncwarp();                                                 // every lane is done reading this buffer
        if (lane == 0) issue(par);
    }
    // the last CTA to finish re-arms the ticket counter
    BFE_TRACE_PT(5, 1);
    __syncthreads();
    BFE_TRACE_PT(5, 2);
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) { counter[0] = 0u; counter[1] = 0u; }
    }
}

// (Two consecutive particles per thread with 16-byte vector loads / stores of inv, R and the outputs -- half the LSU
// instructions -- measured SLOWER, 25.1 vs 23.4 us per 10^6: the random slot reads need the thread count, not fewer instructions.)
__global__ void __launch_bounds__(256)
eof_force_gather_kernel(int64_t n, const int* __restrict__ inv, const double* __restrict__ r_orig,
                        const double2* __restrict__ tmp, double* __restrict__ p0, double* __restrict__ p,
                        double* __restrict__ fr, double* __restrict__ fp, double* __restrict__ fz,
                        double* __restrict__ R) {
    bfe_pdl_wait();
    bfe_pdl_trigger();
    BFE_TRACE_PT(6, 0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2* src = tmp + 3 * (int64_t)__ldg(inv + i);
        const double2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        p0[i] = a.x; p[i] = a.y; fr[i] = b.x; fp[i] = b.y; fz[i] = c.x; R[i] = __ldg(r_orig + i);
    }
    BFE_TRACE_PT(6, 1);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
// option "grid_pct": the persistent grids of the sorted step sized for this share of the SMs, so that the kernels of
// another stream fit beside them (several particle sets in flight)
static inline int bfe_eff_sms(const bfe_eof* h) { int v = h->num_sms * g_bfe_grid_pct / 100; return v < 1 ? 1 : v; }

struct SortWs {
    int* hist; int* cell_start; int* cursor; EofRec* rec; double* r_orig; double2* tmp; int* inv; double* seg;
    int* H;                // [tiles][ncell] of the stable tile sort
};
int g_bfe_sort_stable = 1;     // option "sort_stable": stable, atomic-free tile counting sort (1) / slot claims by global integer atomics (0)

static int64_t tile_rows_for(const bfe_eof* h, int64_t cap) {
    const int64_t a = h->num_sms, b = cap / 8192 + 2;
    return a > b ? a : b;
}

static int sort_workspace(bfe_eof* h, int64_t n, SortWs* ws) {
    const int ncell = h->g.numx * h->g.numy;
    size_t o_hist = 0;
    size_t o_start = align_up(o_hist + sizeof(int) * ncell, 256);
    size_t o_cur = align_up(o_start + sizeof(int) * (ncell + 1), 256);
    size_t o_rec = align_up(o_cur + sizeof(int) * ncell * BFE_CURSOR_STRIDE, 256);
    if (n > h->sort_cap || !h->sort_ws) {
        if (h->sort_ws) { BFE_CUDA(cudaDeviceSynchronize()); BFE_CUDA(cudaFree(h->sort_ws)); h->sort_ws = nullptr; }
        int64_t cap = (n + n / 8 + 1024 + 15) / 16 * 16;      // multiple of 16: keeps every sub-array 16-B aligned
        // per particle: 64-B record, R (8 B), 48-B force slot, inverse permutation (4 B); then the segment
        // tiles (512 B per cell and per 128-record task)
        BFE_CUDA(cudaMalloc(&h->sort_ws, o_rec + (sizeof(EofRec) + sizeof(double) + 48 + sizeof(int)) * (size_t)cap +
                                             512 * ((size_t)ncell + (size_t)cap / SegSum::TASK + 2) +
                                             sizeof(int) * (size_t)ncell * (size_t)tile_rows_for(h, cap) + 256));
        BFE_CUDA(cudaMemset(h->sort_ws, 0, o_rec));
        BFE_CUDA(cudaDeviceSynchronize());
        h->sort_cap = cap;
    }
    char* b = (char*)h->sort_ws;
    ws->hist = (int*)(b + o_hist); ws->cell_start = (int*)(b + o_start); ws->cursor = (int*)(b + o_cur);
    ws->rec = (EofRec*)(b + o_rec);
    ws->r_orig = (double*)(b + o_rec + sizeof(EofRec) * (size_t)h->sort_cap);
    ws->tmp = (double2*)(b + o_rec + (sizeof(EofRec) + sizeof(double)) * (size_t)h->sort_cap);
    ws->inv = (int*)(b + o_rec + (sizeof(EofRec) + sizeof(double) + 48) * (size_t)h->sort_cap);
    ws->seg = (double*)(b + o_rec + (sizeof(EofRec) + sizeof(double) + 48 + sizeof(int)) * (size_t)h->sort_cap);
    ws->H = (int*)(b + align_up(o_rec + (sizeof(EofRec) + sizeof(double) + 48 + sizeof(int)) * (size_t)h->sort_cap +
                                512 * ((size_t)ncell + (size_t)h->sort_cap / SegSum::TASK + 2), 256));
    return BFE_OK;
}

// counting sort of the particle set by table cell into the handle's workspace
extern "C" int bfe_eof_prepare(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                               const double* mass, void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (n > 0 && (!x || !y || !z)) return BFE_ERR_ARG;
    if (n >= (int64_t)1 << 31) return BFE_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    SortWs ws;
    h->prepared_n = -1;
    int rc = sort_workspace(h, n, &ws);
    if (rc != BFE_OK) return rc;
    const int ncell = h->g.numx * h->g.numy;
    if (g_bfe_sort_stable && n > 0) {
        // tiles of 1024 U consecutive particles, about one per SM, U <= 8 (ids fit 13 bits; rows fit the workspace)
        int64_t U = (n + (int64_t)bfe_eff_sms(h) * 1024 - 1) / ((int64_t)bfe_eff_sms(h) * 1024);
        U = U < 1 ? 1 : (U > 8 ? 8 : U);
        const int tile = (int)(U * 1024);
        const int ntile = (int)((n + tile - 1) / tile);
        const int per = (ncell + 1023) / 1024;
        const size_t ss_h = sizeof(int) * (size_t)per * 1024;
        const size_t ss_s = sizeof(int) * (size_t)per * 1024 + sizeof(unsigned short) * (size_t)tile + 16;
        if (ss_s <= 200 * 1024 && (int64_t)ntile <= tile_rows_for(h, h->sort_cap)) {
            if (ss_h > 48 * 1024) {
                BFE_CUDA(cudaFuncSetAttribute(eof_tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss_h));
            }
            if (ss_s > 48 * 1024)
                BFE_CUDA(cudaFuncSetAttribute(eof_tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss_s));
            int kt = bfe_kt_begin("eof_cell_hist_kernel", stream);
            BFE_CUDA(bfe_launch(eof_tile_hist_kernel, dim3(ntile), dim3(1024), ss_h, stream, nullptr, 0, h->g, ncell, n, tile, x, y,
                                z, ws.H, ws.inv));
            bfe_kt_end(kt, stream);
            kt = bfe_kt_begin("eof_tile_colscan_kernel", stream);
            BFE_CUDA(bfe_launch(eof_tile_colscan_kernel, dim3((ncell + BFE_COLSCAN_CW - 1) / BFE_COLSCAN_CW), dim3(1024), 0, stream, nullptr, 0, ncell,
                                ntile, ws.H, ws.hist, ws.cell_start, h->counter));
            bfe_kt_end(kt, stream);
            BFE_LAUNCH_CHECK("eof_tile_hist_kernel");
            bfe_count_launch(1);
            kt = bfe_kt_begin("eof_cell_scatter_kernel", stream);
            BFE_CUDA(bfe_launch(eof_tile_scatter_kernel, dim3(ntile), dim3(1024), ss_s, stream, nullptr, 0, h->g, ncell, n, tile, x,
                                y, z, mass, (const int*)ws.cell_start, (const int*)ws.H, ws.rec, ws.inv, ws.r_orig));
            bfe_kt_end(kt, stream);
            BFE_LAUNCH_CHECK("eof_tile_scatter_kernel");
            h->prepared_n = n;
            h->prepared_has_mass = mass ? 1 : 0;
            return BFE_OK;
        }
    }
    int grid = (int)((n + 2047) / 2048);
    if (grid > bfe_eff_sms(h)) grid = bfe_eff_sms(h);        // one 1024-thread CTA per SM: the per-CTA merge is paid once per SM
    if (grid < 1) grid = 1;
    {
        const int per = (ncell + 1023) / 1024;
        const size_t ss = sizeof(int) * (size_t)per * 1024;
        if (ss > 200 * 1024) return BFE_ERR_UNSUPPORTED;
        if (ss > 48 * 1024)
            BFE_CUDA(cudaFuncSetAttribute(eof_cell_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss));
        const int kt = bfe_kt_begin("eof_cell_hist_kernel", stream);
        BFE_CUDA(bfe_launch(eof_cell_hist_kernel, dim3(grid), dim3(1024), ss, stream, nullptr, 0, h->g, ncell, n, x, y, z,
                            ws.hist, ws.cell_start, ws.cursor, h->counter, ws.inv));
        bfe_kt_end(kt, stream);
    }
    BFE_LAUNCH_CHECK("eof_cell_hist_kernel");
    int g2 = (int)((n + 511) / 512);
    if (g2 > bfe_eff_sms(h) * 8) g2 = bfe_eff_sms(h) * 8;
    if (g2 < 1) g2 = 1;
    const int kt2 = bfe_kt_begin("eof_cell_scatter_kernel", stream);
    BFE_CUDA(bfe_launch(eof_cell_scatter_kernel, dim3(g2), dim3(256), 0, stream, nullptr, 0, h->g, n, x, y, z, mass,
                        ws.cell_start, ws.cursor, ws.rec, ws.inv, ws.r_orig));
    bfe_kt_end(kt2, stream);
    BFE_LAUNCH_CHECK("eof_cell_scatter_kernel");
    h->prepared_n = n;
    h->prepared_has_mass = (mass || n == 0) ? 1 : 0;
    return BFE_OK;
}

extern "C" int bfe_eof_accumulate_prepared(bfe_eof* h, double* cos_out, double* sin_out, void* stream_) {
    if (!h || !cos_out || !sin_out) return BFE_ERR_ARG;
    if (h->prepared_n < 0 || !h->prepared_has_mass) return BFE_ERR_STATE;
    if (h->g.mmax > 6 || h->nch_pad > 256) return BFE_ERR_UNSUPPORTED;
    cudaStream_t stream = (cudaStream_t)stream_;
    SortWs ws;
    int rc = sort_workspace(h, h->prepared_n, &ws);
    if (rc != BFE_OK) return rc;
    const int64_t n = h->prepared_n;
    const int ncell = h->g.numx * h->g.numy;
    {
        int64_t nblk = (n + SegSum::TASK * SegSum::NW - 1) / (SegSum::TASK * SegSum::NW);
        int grid = (int)(nblk < (int64_t)bfe_eff_sms(h) * 4 ? nblk : (int64_t)bfe_eff_sms(h) * 4);
        if (grid < 1) grid = 1;
        const int kt = bfe_kt_begin("eof_segsum_kernel", stream);
        BFE_CUDA(bfe_launch(eof_segsum_kernel, dim3(grid), dim3(256), 0, stream, nullptr, 0, n, ws.rec, ws.seg, h->counter));
        bfe_kt_end(kt, stream);
        BFE_LAUNCH_CHECK("eof_segsum_kernel");
    }
    {
        int grid = (ncell + 7) / 8;
        if (grid > bfe_eff_sms(h) * 2) grid = bfe_eff_sms(h) * 2;      // <= 64 reduce groups of 16 CTAs
        if (grid > h->max_ctas) grid = h->max_ctas;
        if (grid > 64 * 16) grid = 64 * 16;
        const int kt = bfe_kt_begin("eof_node_contract_kernel", stream);
        BFE_CUDA(bfe_launch(eof_node_contract_kernel, dim3(grid), dim3(256), 0, stream, h->t_acc,
                            (size_t)h->g.nnode * h->nch_pad * sizeof(double), h->g, h->t_acc, h->nch, h->nch_pad, ncell,
                            ws.cell_start, ws.seg, h->partial, h->counter, cos_out, sin_out));
        bfe_kt_end(kt, stream);
    }
    BFE_LAUNCH_CHECK("eof_node_contract_kernel");
    return BFE_OK;
}

extern "C" int bfe_eof_force_prepared(bfe_eof* h, double* p0, double* p, double* fr, double* fp, double* fz,
                                      double* R, void* stream_) {
    if (!h) return BFE_ERR_ARG;
    if (h->prepared_n < 0 || !h->contracted) return BFE_ERR_STATE;
    if (h->g.mmax > 6) return BFE_ERR_UNSUPPORTED;
    const int64_t n = h->prepared_n;
    if (n == 0) return BFE_OK;
    if (!p0 || !p || !fr || !fp || !fz || !R) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    SortWs ws;
    int rc = sort_workspace(h, n, &ws);
    if (rc != BFE_OK) return rc;
    int64_t need = (n + 127) / 128, cap = (int64_t)bfe_eff_sms(h) * 16;
    int grid = (int)(need < cap ? need : cap);
    if (g_bfe_force_mma) {
        // one resident wave (occupancy queried once: registers and 16.4 kB of record buffers per CTA)
        static int occ = 0;
        if (!occ) {
            BFE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, eof_force_sorted_mma_kernel<6>, 128, 0));
            if (occ < 1) occ = 1;
        }
        int64_t need_m = (n + 511) / 512, cap_m = (int64_t)bfe_eff_sms(h) * occ;
        grid = (int)(need_m < cap_m ? need_m : cap_m);
        const int kt = bfe_kt_begin("eof_force_sorted_mma_kernel", stream);
        BFE_CUDA(bfe_launch(eof_force_sorted_mma_kernel<6>, dim3(grid), dim3(128), 0, stream, h->g_con,
                            (size_t)h->g.nnode * h->gstride * sizeof(double), h->g, h->g_con, h->gstride, n, ws.rec, ws.tmp,
                            h->counter));
        bfe_kt_end(kt, stream);
        BFE_LAUNCH_CHECK("eof_force_sorted_mma_kernel");
    } else {
        const int kt = bfe_kt_begin("eof_force_sorted_kernel", stream);
        BFE_CUDA(bfe_launch(eof_force_sorted_kernel<6>, dim3(grid), dim3(128), 0, stream, h->g_con,
                            (size_t)h->g.nnode * h->gstride * sizeof(double), h->g, h->g_con, h->gstride, n, ws.rec, ws.tmp));
        bfe_kt_end(kt, stream);
        BFE_LAUNCH_CHECK("eof_force_sorted_kernel");
    }
    int64_t need2 = (n + 255) / 256, cap2 = (int64_t)bfe_eff_sms(h) * 8;
    const int kt3 = bfe_kt_begin("eof_force_gather_kernel", stream);
    BFE_CUDA(bfe_launch(eof_force_gather_kernel, dim3((int)(need2 < cap2 ? need2 : cap2)), dim3(256), 0, stream, nullptr, 0,
                        n, ws.inv, ws.r_orig, ws.tmp, p0, p, fr, fp, fz, R));
    bfe_kt_end(kt3, stream);
    BFE_LAUNCH_CHECK("eof_force_gather_kernel");
    return BFE_OK;
}

int bfe_eof_accumulate_sorted(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                              const double* mass, double* cos_out, double* sin_out, cudaStream_t stream) {
    int rc = bfe_eof_prepare(h, n, x, y, z, mass, (void*)stream);
    if (rc != BFE_OK) return rc;
    return bfe_eof_accumulate_prepared(h, cos_out, sin_out, (void*)stream);
}

int bfe_eof_force_sorted(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                         double* p0, double* p, double* fr, double* fp, double* fz, double* R, cudaStream_t stream) {
    int rc = bfe_eof_prepare(h, n, x, y, z, nullptr, (void*)stream);
    if (rc != BFE_OK) return rc;
    return bfe_eof_force_prepared(h, p0, p, fr, fp, fz, R, (void*)stream);
}
