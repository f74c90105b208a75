// bfe_sl_sort.cu -- radial-bin-sorted SL accumulation (spheresl.compute_coefficients_solitary,
// spheresl.py:567-656), the SL counterpart of bfe_sort.cu.
//
// The direct kernel (bfe_sl.cu) reads two radial-node rows (2 x (lmax+1) nmax doubles, ~2 kB) per
// particle from L2.  Linear interpolation is linear in the table, so for all particles of one radial
// interval i
//     expcoef[k,n] += D1[k] E[i][l(k)][n] + D2[k] E[i+1][l(k)][n],
//     D1[k] = sum_p w_k(p) x1(p),  D2[k] = sum_p w_k(p) x2(p),
//     w_k(p) = -4 pi m_p P0(p) f_lm P_l^m(cos theta_p) {1 | cos m phi_p | sin m phi_p}.
// Particles are counting-sorted by radial interval (integer keys and cursors only), a warp sums D over a
// run of equal intervals in registers and reads the two node rows once per RUN (TMA bulk copy,
// prefetched when the run opens).  No floating-point atomics.
//
//   sl_bin_hist_kernel    : radial interval per particle -> histogram; last CTA scans it
//   sl_bin_scatter_kernel : 64-B record {x1, x2, W = -4 pi m P0, cos theta, cos phi, sin phi, -, bin:perm}
//   sl_deposit_kernel     : runs -> D -> coefficient partials
//   sl_sorted_reduce_kernel : column sums of the per-CTA partials
#include "bfe_device.cuh"
#include "bfe_sortcore.cuh"

#define BFE_SL_MAX_GRANULES 131072      // granules of the static deposit split: <= 128 blocks of 1024

struct __align__(16) SlRec {
    double x1, x2;               // linear weights of nodes i, i+1 (spheresl.py:327-328)
    double W;                    // -4 pi m (x1 p0[i] + x2 p0[i+1])
    double costh;
    double c1, s1;               // cos phi, sin phi
    double pad;
    unsigned long long binperm;  // (i << 32) | original particle index
};

struct SlPrep {
    int i;
    double x1, x2, W, costh, c1, s1;
};

__device__ __forceinline__ SlPrep bfe_sl_prep(const SlGeom& g, const double* __restrict__ xi,
                                              const double* __restrict__ p0tab, double px, double py, double pz,
                                              double pm) {
    SlPrep o;
    double r2 = BFE_ADD(BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py)), BFE_MUL(pz, pz));   // spheresl.py:610
    double r = fmax(sqrt(r2), 1.0e-10);                                                // 611
    o.costh = BFE_DIV(pz, r);                                                          // 612
    bfe_cossin_phi(px, py, o.c1, o.s1);                                                // 613
    SlBin b = bfe_sl_bin(g, xi, r);                                                    // 309-328
    o.i = b.i; o.x1 = b.x1; o.x2 = b.x2;
    double P0 = b.x1 * __ldg(p0tab + b.i) + b.x2 * __ldg(p0tab + b.i + 1);
    o.W = BFE_FOURPI_NEG * pm * P0;
    return o;
}

__global__ void __launch_bounds__(1024)
sl_bin_hist_kernel(SlGeom g, const double* __restrict__ xi, int nbin, int64_t n, const double* __restrict__ x,
                   const double* __restrict__ y, const double* __restrict__ z, int* __restrict__ hist,
                   int* __restrict__ bin_start, int* __restrict__ cursor, unsigned int* __restrict__ counter) {
    extern __shared__ int s_hist[];
    __shared__ int s_wsum[32];
    __shared__ bool s_last;
    for (int c = threadIdx.x; c < nbin; c += blockDim.x) s_hist[c] = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nbin; c += gridDim.x * blockDim.x)
        cursor[(size_t)c * BFE_CURSOR_STRIDE] = 0;      // slot-claim counters for the scatter kernel
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        double r2 = BFE_ADD(BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py)), BFE_MUL(pz, pz));
        double r = fmax(sqrt(r2), 1.0e-10);
        SlBin b = bfe_sl_bin(g, xi, r);
        atomicAdd(&s_hist[b.i], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nbin; c += blockDim.x) {
        int v = s_hist[c];
        if (v) atomicAdd(&hist[c], v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        bfe_block_scan_cells(nbin, hist, bin_start, s_hist, s_wsum);
        if (threadIdx.x == 0) *counter = 0u;
    }
}

__global__ void __launch_bounds__(256)
sl_bin_scatter_kernel(SlGeom g, const double* __restrict__ xi, const double* __restrict__ p0tab, int64_t n,
                      const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                      const double* __restrict__ mass, const int* __restrict__ bin_start, int* __restrict__ cursor,
                      SlRec* __restrict__ rec) {
    constexpr int U = 2;
    for (int64_t base = (int64_t)blockIdx.x * (256 * U); base < n; base += (int64_t)gridDim.x * (256 * U)) {
        SlPrep pr[U];
        int pos[U];
        int64_t idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            idx[u] = base + u * 256 + threadIdx.x;
            const bool on = idx[u] < n;
            double px = on ? __ldg(x + idx[u]) : 1.0, py = on ? __ldg(y + idx[u]) : 0.0, pz = on ? __ldg(z + idx[u]) : 0.0;
            double pm = on ? __ldg(mass + idx[u]) : 0.0;
            pr[u] = bfe_sl_prep(g, xi, p0tab, px, py, pz, pm);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) pos[u] = (idx[u] < n) ? (__ldg(bin_start + pr[u].i) + atomicAdd(&cursor[(size_t)pr[u].i * BFE_CURSOR_STRIDE], 1)) : 0;   // integer slot claim
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (idx[u] < n) {
                unsigned long long bp = ((unsigned long long)(unsigned int)pr[u].i << 32) |
                                        (unsigned long long)(unsigned int)idx[u];
                char* dst = reinterpret_cast<char*>(rec + pos[u]);       // two full-sector stores per record
                bfe_st256(dst, pr[u].x1, pr[u].x2, pr[u].W, pr[u].costh);
                bfe_st256(dst + 32, pr[u].c1, pr[u].s1, 0.0, __longlong_as_double((long long)bp));
            }
        }
    }
}

// ---------------------------------------------------------------------------
// STABLE radial-bin sort without global atomics (option "sort_stable", default 1): the tile counting sort of bfe_sort.cu
// (eof_tile_hist_kernel / eof_tile_colscan_kernel / eof_tile_scatter_kernel) with the radial interval as the key.  Together
// with the static task assignment of sl_deposit_kernel it makes the sorted SL accumulation bit-reproducible.
//   sl_tile_hist_kernel    : one CTA per tile of 1024 U consecutive particles: interval per particle (FP64, as the record
//                            needs it), shared-memory integer counters -> row H[tile][bin]; the interval is kept in binid[]
//   (bfe_tile_colscan)     : H[tile][bin] -> members of the bin in earlier tiles; bin_start
//   sl_tile_scatter_kernel : shared-memory counting sort of the tile's ids by bin, rank = members of the own (tile, bin)
//                            group with a smaller id; position = bin_start[bin] + H[tile][bin] + rank; record stored
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sl_tile_hist_kernel(SlGeom g, const double* __restrict__ xi, int nbin, int64_t n, int tile, const double* __restrict__ x,
                    const double* __restrict__ y, const double* __restrict__ z, int* __restrict__ H, int* __restrict__ binid) {
    extern __shared__ int s_hist[];
    for (int c = threadIdx.x; c < nbin; c += blockDim.x) s_hist[c] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * tile;
    for (int o = threadIdx.x; o < tile; o += 1024) {
        const int64_t i = base + o;
        if (i >= n) break;
        const double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        const double r2 = BFE_ADD(BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py)), BFE_MUL(pz, pz));
        const double r = fmax(sqrt(r2), 1.0e-10);
        const SlBin b = bfe_sl_bin(g, xi, r);
        atomicAdd(&s_hist[b.i], 1);                       // shared-memory integer count
        binid[i] = b.i;
    }
    __syncthreads();
    int* row = H + (size_t)blockIdx.x * nbin;
    for (int c = threadIdx.x; c < nbin; c += blockDim.x) row[c] = s_hist[c];
}

__global__ void __launch_bounds__(1024)
sl_tile_scatter_kernel(SlGeom g, const double* __restrict__ xi, const double* __restrict__ p0tab, int nbin, int64_t n, int tile,
                       const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                       const double* __restrict__ mass, const int* __restrict__ bin_start, const int* __restrict__ H,
                       SlRec* __restrict__ rec, const int* __restrict__ binid) {
    extern __shared__ int s_mem[];
    __shared__ int s_wsum[32];
    int* s_cnt = s_mem;                                   // [per * 1024]: counts -> starts -> group ends (as in eof_tile_scatter_kernel)
    const int per = (nbin + 1023) / 1024;
    unsigned short* s_list = reinterpret_cast<unsigned short*>(s_mem + per * 1024);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < per * 1024; c += 1024) s_cnt[c] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * tile;
    const int nu = tile >> 10;                            // particles per thread (<= 8)
    int bin[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        bin[u] = -1;
        if (u < nu) {
            const int64_t i = base + u * 1024 + tid;
            if (i < n) { bin[u] = binid[i]; atomicAdd(&s_cnt[bin[u]], 1); }
        }
    }
    __syncthreads();
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
        for (int k = 0; k < per; ++k) { const int hcnt = s_cnt[lo + k]; s_cnt[lo + k] = run; run += hcnt; }
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 8; ++u)
        if (bin[u] >= 0) s_list[atomicAdd(&s_cnt[bin[u]], 1)] = (unsigned short)(u * 1024 + tid);
    __syncthreads();
    auto bin_at = [&](int u) {
        int c = bin[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) c = (u == k) ? bin[k] : c;
        return c;
    };
#pragma unroll 1
    for (int u = 0; u < nu; ++u) {
        const int c = bin_at(u);
        if (c < 0) continue;
        const int64_t idx = base + u * 1024 + tid;
        const double px = __ldg(x + idx), py = __ldg(y + idx), pz = __ldg(z + idx), pm = __ldg(mass + idx);
        const int gbase = __ldg(bin_start + c) + __ldg(H + (size_t)blockIdx.x * nbin + c);
        const int lo = c ? s_cnt[c - 1] : 0, hi = s_cnt[c];
        const int me = u * 1024 + tid;
        int rank = 0;
        int k = lo;
        for (; (k & 3) && k < hi; ++k) rank += (s_list[k] < me) ? 1 : 0;
        const unsigned int K = (unsigned int)me * 0x00010001u + 0x7fff7fffu;      // see eof_tile_scatter_kernel
        unsigned int acc = 0u;
        for (; k + 3 < hi; k += 4) {
            const uint2 w = *reinterpret_cast<const uint2*>(s_list + k);
            acc += (((K - w.x) >> 15) & 0x00010001u) + (((K - w.y) >> 15) & 0x00010001u);
        }
        rank += (int)((acc & 0xffffu) + (acc >> 16));
        for (; k < hi; ++k) rank += (s_list[k] < me) ? 1 : 0;
        const int pos = gbase + rank;
        const SlPrep pr = bfe_sl_prep(g, xi, p0tab, px, py, pz, pm);        // pr.i == c
        const unsigned long long bp = ((unsigned long long)(unsigned int)c << 32) | (unsigned long long)(unsigned int)idx;
        char* dst = reinterpret_cast<char*>(rec + pos);
        bfe_st256(dst, pr.x1, pr.x2, pr.W, pr.costh);
        bfe_st256(dst + 32, pr.c1, pr.s1, 0.0, __longlong_as_double((long long)bp));
    }
}

// ---------------------------------------------------------------------------
// Static, balanced work split for the deposit kernel (with the stable sort; option sort_stable).  The sorted array is cut
// into granules of G records; the cost of a granule is modelled as G + FC * (radial intervals it spans, <= G) -- every run of
// equal intervals ends in a flush that costs about FC records' worth of time -- and warp w of the grid gets the contiguous
// granules whose cumulative cost lies in [w, w + 1) * total / W.  The split is a function of the sorted data alone, so the
// per-warp sums, the per-CTA partials and the coefficients are bit-reproducible, and a warp walks ONE contiguous range:
// runs that continue across granules are not flushed in between (the dynamic queue flushed at the end of every
// 128-record task).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
sl_granule_cost_kernel(int64_t n, int gran, int ngran, int flush_cost, const SlRec* __restrict__ rec,
                       unsigned int* __restrict__ lprefix, unsigned int* __restrict__ btot) {
    // thread per granule; exclusive prefix inside the block of 1024 granules -> lprefix, block total -> btot (the deposit
    // kernel scans the <= 128 block totals itself)
    __shared__ unsigned int s_w[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x * 1024 + tid;
    unsigned int cost = 0u;
    if (g < ngran) {
        const int64_t a = (int64_t)g * gran, b = min(n, a + gran) - 1;
        const int ba = (int)((unsigned long long)__double_as_longlong(__ldg(reinterpret_cast<const double*>(rec + a) + 7)) >> 32);
        const int bb = (int)((unsigned long long)__double_as_longlong(__ldg(reinterpret_cast<const double*>(rec + b) + 7)) >> 32);
        int span = bb - ba + 1;
        if (span < 1) span = 1;
        if (span > gran) span = gran;
        cost = (unsigned int)((int)(b - a + 1) + flush_cost * span);
    }
    unsigned int incl = cost;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned int w = s_w[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, w, off); if (lane >= off) w += v; }
        s_w[lane] = w;
    }
    __syncthreads();
    if (g < ngran) lprefix[g] = incl - cost + (warp > 0 ? s_w[warp - 1] : 0u);
    if (tid == 1023) btot[blockIdx.x] = s_w[31];
}

// ---------------------------------------------------------------------------
// deposit: a warp owns TASK consecutive sorted records (dynamic task queue).
//   expand (lanes = records): P_l^m recurrence (explicitly rounded, as bfe_legendre), cos/sin(m phi),
//     w_k for all (lmax+1)^2 rows into the warp's shared-memory slab, plus x1, x2;
//   sum   (lanes = rows k, two per lane): D1[k] += w_k x1, D2[k] += w_k x2 over the run;
//   flush (cell change, warp-uniform): lane j = lane + 32 c adds D1[k(j)] E[i][..] + D2[k(j)] E[i+1][..]
//     to its register accumulators; the two node rows were fetched by one TMA bulk copy when the run opened.
// ---------------------------------------------------------------------------
#ifdef BFE_PROFILE_DEPOSIT
__device__ long long* g_sl_dbg = nullptr;
extern "C" int bfe_sl_debug_set(long long* p) { return (int)cudaMemcpyToSymbol(g_sl_dbg, &p, sizeof(p)); }
#define SDBG_DECL long long sd_t0 = clock64(), sd_tk = 0, sd_expand = 0, sd_sum = 0, sd_flush = 0, sd_nflush = 0, sd_ntask = 0, sd_wait = 0
#define SDBG_TICK() (sd_tk = clock64())
#define SDBG_ADD(var) do { long long t_ = clock64(); var += t_ - sd_tk; sd_tk = t_; } while (0)
#else
#define SDBG_DECL
#define SDBG_TICK()
#define SDBG_ADD(var)
#endif

template <int LCAP, int KC>
__global__ void __launch_bounds__(128, 3)
sl_deposit_kernel(SlGeom g, const double* __restrict__ e_node, const double* __restrict__ fac, int no_odd,
                  int64_t n, const SlRec* __restrict__ rec, double* __restrict__ partial,
                  unsigned int* __restrict__ counter, int use_tma, int static_tasks, int gran, int ngran,
                  const unsigned int* __restrict__ gprefix) {
    constexpr int TASK = 128;
    constexpr int NW = 4;                          // warps per CTA
    constexpr int NROWCAP = (LCAP + 1) * (LCAP + 1);
    constexpr int RS = 33;                         // slab row stride: lanes = rows read one column conflict-free
    constexpr int SLAB = (NROWCAP + 2) * RS;
    constexpr int TB = 2 * (LCAP + 1) * 32;        // table rows of the open run: 2 nodes x ln (<= (LCAP+1)*32) doubles
    extern __shared__ __align__(128) unsigned char s_raw[];
    double* s_tb = reinterpret_cast<double*>(s_raw);                       // [NW][TB]
    double* s_slab = s_tb + NW * TB;                                       // [NW][SLAB]
    double* s_D = s_slab + NW * SLAB;                                      // [NW][2][64]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_D + NW * 128);
    int* s_idx = reinterpret_cast<int*>(s_bar + NW);                      // [32*KC] coefficient j -> (row k << 16) | (l*nmax+n)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nrow = g.nrow, ln = g.ln, nmax = g.nmax;
    const int ncoef = nrow * nmax;
    for (int j = tid; j < 32 * KC; j += blockDim.x) {
        int v = 0;
        if (j < ncoef) {
            const int k = j / nmax, nn = j - k * nmax;
            int l = (int)sqrtf((float)k);
            if ((l + 1) * (l + 1) <= k) ++l;
            if (l * l > k) --l;
            v = (k << 16) | (l * nmax + nn);
        }
        s_idx[j] = v;
    }
    const int kA = lane, kB = lane + 32;
    const bool a_on = kA < nrow, b_on = kB < nrow;
    double acc[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) acc[c] = 0.0;

    double* val = s_slab + warp * SLAB;            // val[k * RS + p]; rows nrow, nrow+1 hold x1, x2
    double* D = s_D + warp * 128;
    double* tb = s_tb + warp * TB;
    const unsigned int bar = bfe_smem_u32(s_bar + warp);
    const unsigned int tb_u32 = bfe_smem_u32(tb);
    const unsigned int rowbytes = 2u * (unsigned int)ln * 8u;
    unsigned int bar_parity = 0;
    if (lane == 0) bfe_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    unsigned int* task_counter = counter + 1;
    const int64_t ntasks = (n + TASK - 1) / TASK;
    SDBG_DECL;
    // static_tasks (with the stable sort): this warp's contiguous range of granules from the cost prefix (sl_granule_cost_kernel)
    // -- a function of the sorted data alone, so the sums are bit-reproducible; the dynamic queue made the per-CTA partials depend
    // on the schedule (coefficients reproducible to ~1e-16 only)
    int64_t s_t0 = 0, s_cnt = 0;
    if (static_tasks) {
        // gprefix = [lprefix (ngran) | btot (<= 128 blocks of 1024 granules)]: block prefixes by one warp scan, four per lane
        const unsigned int* btot = gprefix + BFE_SL_MAX_GRANULES;
        const int nblkg = (ngran + 1023) >> 10;
        unsigned int bt[4], bp[4];
        unsigned int mine = 0u;
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int bidx = lane * 4 + q; bt[q] = bidx < nblkg ? __ldg(btot + bidx) : 0u; mine += bt[q]; }
        unsigned int incl = mine;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += v; }
        const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
        bp[0] = incl - mine; bp[1] = bp[0] + bt[0]; bp[2] = bp[1] + bt[1]; bp[3] = bp[2] + bt[2];     // exclusive block prefixes
        const unsigned long long gw = (unsigned long long)blockIdx.x * NW + warp, W = (unsigned long long)gridDim.x * NW;
        const unsigned int lo_c = (unsigned int)(total * gw / W), hi_c = (unsigned int)(total * (gw + 1) / W);
        auto lower = [&](unsigned int c) {                 // first granule g with (global) prefix[g] >= c
            // block: the last one whose prefix is <= c (blocks are non-empty, prefixes strictly increase)
            int cnt = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) cnt += (lane * 4 + q < nblkg && bp[q] <= c) ? 1 : 0;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
            const int blk = cnt > 0 ? cnt - 1 : 0;
            unsigned int base = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) { const unsigned int v = __shfl_sync(0xffffffffu, bp[q], blk >> 2); if ((blk & 3) == q) base = v; }
            const unsigned int cl = c - base;
            int a = blk << 10, b = min(ngran, a + 1024);
            while (a < b) { const int mid = (a + b) >> 1; if (__ldg(gprefix + mid) < cl) a = mid + 1; else b = mid; }
            return a;
        };
        const int glo = lower(lo_c), ghi = (gw + 1 == W) ? ngran : lower(hi_c);
        s_t0 = (int64_t)glo * gran;
        const int64_t e = (int64_t)ghi * gran < n ? (int64_t)ghi * gran : n;
        s_cnt = e > s_t0 ? e - s_t0 : 0;
    }
    for (;;) {
        int64_t t0;
        int tcnt;
        if (static_tasks) {
            if (s_cnt <= 0) break;
            t0 = s_t0; tcnt = (int)s_cnt;
            s_cnt = 0;                                      // one pass over the whole range
        } else {
            int64_t task = 0;
            if (lane == 0) task = (int64_t)atomicAdd(task_counter, 1u);
            task = __shfl_sync(0xffffffffu, task, 0);
            if (task >= ntasks) break;
            t0 = task * TASK;
            tcnt = (int)((n - t0) < TASK ? (n - t0) : TASK);
        }
#ifdef BFE_PROFILE_DEPOSIT
        sd_ntask++;
#endif
        double dA1 = 0.0, dA2 = 0.0, dB1 = 0.0, dB2 = 0.0;
        int cur = -1;
        double2 ra, rb, rc, rd;
        {
            const bool on = lane < ((tcnt < 32) ? tcnt : 32);
            const double2* src = reinterpret_cast<const double2*>(rec + t0 + (on ? lane : 0));
            ra = __ldg(src); rb = __ldg(src + 1); rc = __ldg(src + 2); rd = __ldg(src + 3);
        }
        for (int b0 = 0; b0 < tcnt; b0 += 32) {
            const int bcnt = (tcnt - b0) < 32 ? (tcnt - b0) : 32;
            __syncwarp();
            SDBG_TICK();
            int mybin = -3;
            {
                const bool on = lane < bcnt;
                const double W = on ? rb.x : 0.0;                  // zero weight parks harmless values
#ifdef BFE_PROFILE_DEPOSIT
                if (__double_as_longlong(W) == 0x7ff8dead00000000ll) sd_wait = 1;
                SDBG_ADD(sd_wait);
#endif
                const double x = rb.y;
                if (on) mybin = (int)((unsigned long long)__double_as_longlong(rd.y) >> 32);
                val[nrow * RS + lane] = ra.x;
                val[(nrow + 1) * RS + lane] = ra.y;
                // P_l^m columns (m outer), spheresl.py:664-700.  Accumulation uses no derivative, so the
                // recurrence is evaluated with fused multiply-adds and reciprocal constants (<= few ulp)
                const double somx2 = sqrt((1.0 - x) * (1.0 + x));
                double pmm = 1.0, fact = 1.0, cm = 1.0, sm = 0.0;
#pragma unroll
                for (int m = 0; m <= LCAP; ++m) {
                    if (m <= g.lmax) {
                        if (m > 0) {
                            pmm = pmm * (-fact * somx2);
                            fact += 2.0;
                            double cn = cm * rc.x - sm * rc.y, sn = sm * rc.x + cm * rc.y;
                            cm = cn; sm = sn;
                        }
                        double pl2 = 0.0, pl1 = pmm;
#pragma unroll
                        for (int l = m; l <= LCAP; ++l) {
                            if (l <= g.lmax) {
                                double P;
                                if (l == m) P = pmm;
                                else {
                                    if (l == m + 1) P = x * (2.0 * m + 1.0) * pl1;
                                    else P = (x * (double)(2 * l - 1) * pl1 - (double)(l + m - 1) * pl2) *
                                             (1.0 / (double)(l - m));
                                    pl2 = pl1; pl1 = P;
                                }
                                const bool skip = no_odd && (l & 1);          // spheresl.py:630-632
                                const double fw = skip ? 0.0 : W * __ldg(fac + l * (g.lmax + 1) + m) * P;
                                if (m == 0) val[(l * l) * RS + lane] = fw;
                                else {
                                    val[(l * l + 2 * m - 1) * RS + lane] = fw * cm;
                                    val[(l * l + 2 * m) * RS + lane] = fw * sm;
                                }
                            }
                        }
                    }
                }
            }
            if (b0 + 32 < tcnt) {
                const bool on = (b0 + 32 + lane) < tcnt;
                const double2* src = reinterpret_cast<const double2*>(rec + t0 + b0 + 32 + (on ? lane : 0));
                ra = __ldg(src); rb = __ldg(src + 1); rc = __ldg(src + 2); rd = __ldg(src + 3);
            }
            int prevbin = __shfl_up_sync(0xffffffffu, mybin, 1);
            if (lane == 0) prevbin = cur;
            unsigned int bmask = __ballot_sync(0xffffffffu, (lane < bcnt) && (mybin != prevbin));
            __syncwarp();
            SDBG_ADD(sd_expand);

#define SL_OPEN_RUN(binv)                                                                                 \
            do {                                                                                          \
                cur = (binv);                                                                             \
                if (use_tma && lane == 0) {                                                               \
                    bfe_mbar_expect_tx(bar, rowbytes);                                                    \
                    bfe_bulk_g2s(tb_u32, e_node + (size_t)cur * ln, rowbytes, bar);                       \
                }                                                                                         \
            } while (0)

#define SL_FLUSH_RUN()                                                                                    \
            do {                                                                                          \
                __syncwarp();                                                                             \
                if (a_on) { D[kA] = dA1; D[64 + kA] = dA2; }                                              \
                if (b_on) { D[kB] = dB1; D[64 + kB] = dB2; }                                              \
                const double* e0_;                                                                        \
                if (use_tma) { bfe_mbar_wait(bar, bar_parity); bar_parity ^= 1u; e0_ = tb; }              \
                else e0_ = e_node + (size_t)cur * ln;                                                     \
                __syncwarp();                                                                             \
                _Pragma("unroll")                                                                         \
                for (int c = 0; c < KC; ++c) {                                                            \
                    const int j_ = lane + 32 * c;                                                         \
                    if (j_ < ncoef) {                                                                     \
                        const int pk_ = s_idx[j_];                                                        \
                        const int k_ = pk_ >> 16, eo_ = pk_ & 0xffff;                                     \
                        acc[c] += D[k_] * e0_[eo_] + D[64 + k_] * e0_[ln + eo_];                          \
                    }                                                                                     \
                }                                                                                         \
                __syncwarp();                                                                             \
                dA1 = 0.0; dA2 = 0.0; dB1 = 0.0; dB2 = 0.0;                                               \
            } while (0)

            const double* wa = val + kA * RS;
            const double* wb = val + (b_on ? kB : 0) * RS;
            const double* x1v = val + nrow * RS;
            const double* x2v = val + (nrow + 1) * RS;
            int p = 0;
            while (p < bcnt) {
                if ((bmask >> p) & 1u) {
                    SDBG_TICK();
                    if (cur >= 0) {
                        SL_FLUSH_RUN();
#ifdef BFE_PROFILE_DEPOSIT
                        sd_nflush++;
                        if (__double_as_longlong(acc[0]) == 0x7ff8dead00000000ll) sd_wait = 1;
#endif
                    }
                    SL_OPEN_RUN(__shfl_sync(0xffffffffu, mybin, p));
                    SDBG_ADD(sd_flush);
                }
                SDBG_TICK();
                const unsigned int later = (p < 31) ? (bmask >> (p + 1)) : 0u;
                const int qend = later ? (p + __ffs(later)) : bcnt;
                if (b_on) {
#pragma unroll 4
                    for (int i = p; i < qend; ++i) {
                        const double u1 = x1v[i], u2 = x2v[i], w0 = wa[i], w1 = wb[i];
                        dA1 = fma(w0, u1, dA1); dA2 = fma(w0, u2, dA2);
                        dB1 = fma(w1, u1, dB1); dB2 = fma(w1, u2, dB2);
                    }
                } else if (a_on) {
#pragma unroll 4
                    for (int i = p; i < qend; ++i) {
                        const double w0 = wa[i];
                        dA1 = fma(w0, x1v[i], dA1);
                        dA2 = fma(w0, x2v[i], dA2);
                    }
                }
#ifdef BFE_PROFILE_DEPOSIT
                if (__double_as_longlong(dA1 + dB1) == 0x7ff8dead00000000ll) sd_wait = 1;
#endif
                SDBG_ADD(sd_sum);
                p = qend;
            }
        }
        if (cur >= 0) SL_FLUSH_RUN();
#undef SL_FLUSH_RUN
#undef SL_OPEN_RUN
    }

#ifdef BFE_PROFILE_DEPOSIT
    if (lane == 0 && g_sl_dbg) {
        long long* d = g_sl_dbg + (blockIdx.x * NW + warp) * 8;
        d[0] = clock64() - sd_t0; d[1] = sd_wait; d[2] = sd_expand; d[3] = sd_sum; d[4] = sd_flush; d[5] = sd_nflush;
        d[6] = sd_ntask; d[7] = 0;
    }
#endif
    // ---- CTA combine over the 4 warps (slab reused), per-CTA partial
    __syncthreads();
#pragma unroll
    for (int c = 0; c < KC; ++c) s_slab[warp * SLAB + c * 32 + lane] = acc[c];
    __syncthreads();
    for (int j = tid; j < ncoef; j += 128) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += s_slab[w * SLAB + j];
        partial[(size_t)blockIdx.x * ncoef + j] = v;
    }
}

// ---------------------------------------------------------------------------
// deposit, register formulation (option sl_deposit_mode = 2; NOT the default: measured 349 -> 405 us per 4e6 particles
// at lmax 4 and 574 -> 549 us at lmax 6 for the whole accumulate, profiles/sl_probe.py -- at 222-253 registers only 8
// warps fit an SM and the serial Legendre recurrences leave the FP64 pipe at 18 %: ncu warps_active 10 %, stalls "wait"
// 1.8 and long scoreboard 2.6 per issue).  ncu of the kernel above (4e6 particles, lmax 6): l1tex 63 %, FP64 pipe
// 23 % -- it is bound by the shared-memory slab (49 stores per record, 4 loads per 4 FMAs in the sum phase), not
// by arithmetic.  Here a LANE keeps the run sums of its own records in registers:
//     d1[r] += (f_lm P_l^m trig)(p) * (W x1)(p),   d2[r] += ... * (W x2)(p)        (one FMA each, no memory traffic)
// for the rows r of the azimuthal orders [MLO, MHI] of its warp -- lmax <= 4: one warp covers all 25 rows
// (50 accumulators); lmax <= 6: the 49 rows are split over two kinds of warps (m = 0..2: 29 rows, m = 3..6: 20 rows),
// each kind pulling every task from its own queue.  Runs are ~n/numr records long, so the sums leave the registers
// rarely: when the radial interval changes (warp-uniform test; records are sorted) the 2 NR x 32 lane sums are
// reduced through a shared-memory slab, contracted with the two node rows of the interval (read from L2, coalesced)
// into the warp's coefficient partials (shared memory), and cleared.  A batch that straddles intervals is taken
// interval by interval with the other lanes' weights zeroed.  No floating-point atomics.
// ---------------------------------------------------------------------------
template <int LCAP, int MLO, int MHI>
struct SlLaneRows {
    __host__ __device__ static constexpr int count() {
        int c = 0;
        for (int m = MLO; m <= MHI; ++m) c += (LCAP - m + 1) * (m ? 2 : 1);
        return c;
    }
};

template <int LCAP, int MLO, int MHI, int NR>
__device__ __forceinline__ void sl_lane_accumulate(const SlGeom& g, const double* __restrict__ fac, int no_odd,
                                                   double wx1, double wx2, double x, double c1, double s1,
                                                   double (&d1)[NR], double (&d2)[NR]) {
    const double somx2 = sqrt((1.0 - x) * (1.0 + x));
    double pmm = 1.0, fact = 1.0, cm = 1.0, sm = 0.0;
    int r = 0;
#pragma unroll
    for (int m = 0; m <= MHI; ++m) {
        if (m > 0) {
            pmm = pmm * (-fact * somx2);
            fact += 2.0;
            const double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
        if (m >= MLO) {
            double pl2 = 0.0, pl1 = pmm;
#pragma unroll
            for (int l = m; l <= LCAP; ++l) {
                double fP = 0.0;
                if (l <= g.lmax) {                                        // warp-uniform
                    double P;
                    if (l == m) P = pmm;
                    else {
                        if (l == m + 1) P = x * (2.0 * m + 1.0) * pl1;
                        else P = (x * (double)(2 * l - 1) * pl1 - (double)(l + m - 1) * pl2) * (1.0 / (double)(l - m));
                        pl2 = pl1; pl1 = P;
                    }
                    const bool skip = no_odd && (l & 1);                  // spheresl.py:630-632
                    fP = skip ? 0.0 : __ldg(fac + l * (g.lmax + 1) + m) * P;
                }
                if (m == 0) {
                    d1[r] = fma(fP, wx1, d1[r]); d2[r] = fma(fP, wx2, d2[r]); ++r;
                } else {
                    const double a = fP * cm, b = fP * sm;
                    d1[r] = fma(a, wx1, d1[r]); d2[r] = fma(a, wx2, d2[r]); ++r;
                    d1[r] = fma(b, wx1, d1[r]); d2[r] = fma(b, wx2, d2[r]); ++r;
                }
            }
        }
    }
}

// body of one warp of the register-formulation kernel for the azimuthal orders [MLO, MHI]
template <int LCAP, int MLO, int MHI>
__device__ __forceinline__ void sl_lane_warp(const SlGeom& g, const double* __restrict__ e_node,
                                             const double* __restrict__ fac, int no_odd, int64_t n,
                                             const SlRec* __restrict__ rec, unsigned int* __restrict__ queue,
                                             double* __restrict__ slab, double* __restrict__ Dk,
                                             double* __restrict__ acc_s, const int* __restrict__ s_idx, int lane) {
    constexpr int NR = SlLaneRows<LCAP, MLO, MHI>::count();
    constexpr int TASK = 1024, RS = 33;
    const int ncoef = g.nrow * g.nmax, ln = g.ln;
    double d1[NR], d2[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) { d1[r] = 0.0; d2[r] = 0.0; }
    int cur = -1;

    auto flush = [&]() {
        // lane sums -> slab[row][lane]; lane q then sums rows q, q+32, ... over the 32 columns
        __syncwarp();
#pragma unroll
        for (int r = 0; r < NR; ++r) { slab[r * RS + lane] = d1[r]; slab[(NR + r) * RS + lane] = d2[r]; d1[r] = 0.0; d2[r] = 0.0; }
        __syncwarp();
        for (int q = lane; q < 2 * NR; q += 32) {
            const double* rowp = slab + q * RS;
            double s0 = 0.0, s1_ = 0.0;
#pragma unroll
            for (int c = 0; c < 32; c += 2) { s0 += rowp[c]; s1_ += rowp[c + 1]; }
            // row index r of this warp's list -> table row k (same enumeration as sl_lane_accumulate)
            int rr = (q < NR) ? q : q - NR, k = 0, cnt = 0;
            for (int m = MLO; m <= MHI; ++m)
                for (int l = m; l <= LCAP; ++l) {
                    const int nrow_lm = m ? 2 : 1;
                    if (rr >= cnt && rr < cnt + nrow_lm) k = (m == 0) ? l * l : l * l + 2 * m - 1 + (rr - cnt);
                    cnt += nrow_lm;
                }
            Dk[(q < NR ? 0 : 64) + k] = s0 + s1_;
        }
        __syncwarp();
        const double* e0 = e_node + (size_t)cur * ln;
        for (int j = lane; j < ncoef; j += 32) {
            const int pk = s_idx[j];
            const int mm = (pk >> 26) & 31;
            if (mm >= MLO && mm <= MHI) {
                const int k = (pk >> 16) & 1023, eo = pk & 0xffff;
                acc_s[j] += Dk[k] * __ldg(e0 + eo) + Dk[64 + k] * __ldg(e0 + ln + eo);
            }
        }
        __syncwarp();
    };

    const int64_t ntasks = (n + TASK - 1) / TASK;
    for (;;) {
        int64_t task = 0;
        if (lane == 0) task = (int64_t)atomicAdd(queue, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntasks) break;
        task = ntasks - 1 - task;                      // sparse outer intervals (short runs, many flushes) first
        const int64_t t0 = task * TASK;
        const int tcnt = (int)((n - t0) < TASK ? (n - t0) : TASK);
        // records two batches ahead are kept in flight (a batch is ~600 cycles of arithmetic, less than a DRAM round trip)
        double2 ra, rb, rc, rd, na, nb, nc, nd;
        {
            const bool on = lane < tcnt;
            const double2* src = reinterpret_cast<const double2*>(rec + t0 + (on ? lane : 0));
            ra = __ldg(src); rb = __ldg(src + 1); rc = __ldg(src + 2); rd = __ldg(src + 3);
            const bool on1 = (32 + lane) < tcnt;
            const double2* src1 = reinterpret_cast<const double2*>(rec + t0 + (on1 ? 32 + lane : 0));
            na = __ldg(src1); nb = __ldg(src1 + 1); nc = __ldg(src1 + 2); nd = __ldg(src1 + 3);
        }
        for (int b0 = 0; b0 < tcnt; b0 += 32) {
            const bool on = (b0 + lane) < tcnt;
            const double x1 = ra.x, x2 = ra.y, W = on ? rb.x : 0.0, x = rb.y, c1 = rc.x, s1 = rc.y;
            const int mybin = on ? (int)((unsigned long long)__double_as_longlong(rd.y) >> 32) : -1;
            ra = na; rb = nb; rc = nc; rd = nd;
            if (b0 + 64 < tcnt) {
                const bool on2 = (b0 + 64 + lane) < tcnt;
                const double2* src = reinterpret_cast<const double2*>(rec + t0 + b0 + 64 + (on2 ? lane : 0));
                na = __ldg(src); nb = __ldg(src + 1); nc = __ldg(src + 2); nd = __ldg(src + 3);
            }
            if (__all_sync(0xffffffffu, !on || mybin == cur)) {
                sl_lane_accumulate<LCAP, MLO, MHI, NR>(g, fac, no_odd, W * x1, W * x2, x, c1, s1, d1, d2);
            } else {
                unsigned int todo = __ballot_sync(0xffffffffu, on);
                while (todo) {
                    const int b = __shfl_sync(0xffffffffu, mybin, __ffs(todo) - 1);
                    if (b != cur) { if (cur >= 0) flush(); cur = b; }
                    const bool mine = on && (mybin == b);
                    const double Wm = mine ? W : 0.0;
                    sl_lane_accumulate<LCAP, MLO, MHI, NR>(g, fac, no_odd, Wm * x1, Wm * x2, x, c1, s1, d1, d2);
                    todo &= ~__ballot_sync(0xffffffffu, mine);
                }
            }
        }
    }
    if (cur >= 0) flush();
}

template <int LCAP, int NSPLIT>
__global__ void __launch_bounds__(128, 2)
sl_deposit_lane_kernel(SlGeom g, const double* __restrict__ e_node, const double* __restrict__ fac, int no_odd,
                       int64_t n, const SlRec* __restrict__ rec, double* __restrict__ partial,
                       unsigned int* __restrict__ counter) {
    constexpr int NW = 4, RS = 33;
    constexpr int MSPLIT = 2;                                        // lmax 6: warps of kind 0 take m = 0..2, kind 1 m = 3..6
    constexpr int NR0 = (NSPLIT == 1) ? SlLaneRows<LCAP, 0, LCAP>::count() : SlLaneRows<LCAP, 0, MSPLIT>::count();
    constexpr int SLAB = 2 * NR0 * RS;                               // kind 0 has the larger row set
    extern __shared__ __align__(16) unsigned char s_raw2[];
    double* s_slab = reinterpret_cast<double*>(s_raw2);              // [NW][SLAB]
    double* s_Dk = s_slab + NW * SLAB;                               // [NW][128]
    double* s_acc = s_Dk + NW * 128;                                 // [NW][ncoef_pad]
    const int ncoef = g.nrow * g.nmax, ncoef_pad = (ncoef + 31) & ~31;
    int* s_idx = reinterpret_cast<int*>(s_acc + NW * ncoef_pad);     // [ncoef_pad]: (m << 26) | (k << 16) | (l*nmax + n)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int j = tid; j < ncoef_pad; j += blockDim.x) {
        int v = 0;
        if (j < ncoef) {
            const int k = j / g.nmax, nn = j - k * g.nmax;
            int l = (int)sqrtf((float)k);
            if ((l + 1) * (l + 1) <= k) ++l;
            if (l * l > k) --l;
            const int m = (k - l * l + 1) / 2;
            v = (m << 26) | (k << 16) | (l * g.nmax + nn);
        }
        s_idx[j] = v;
    }
    for (int j = tid; j < NW * ncoef_pad; j += blockDim.x) s_acc[j] = 0.0;
    for (int j = tid; j < NW * 128; j += blockDim.x) s_Dk[j] = 0.0;
    __syncthreads();
    double* slab = s_slab + warp * SLAB;
    double* Dk = s_Dk + warp * 128;
    double* acc_s = s_acc + warp * ncoef_pad;
    if (NSPLIT == 1) {
        sl_lane_warp<LCAP, 0, LCAP>(g, e_node, fac, no_odd, n, rec, counter + 1, slab, Dk, acc_s, s_idx, lane);
    } else if ((warp & 1) == 0) {
        sl_lane_warp<LCAP, 0, MSPLIT>(g, e_node, fac, no_odd, n, rec, counter + 1, slab, Dk, acc_s, s_idx, lane);
    } else {
        sl_lane_warp<LCAP, MSPLIT + 1, LCAP>(g, e_node, fac, no_odd, n, rec, counter + 2, slab, Dk, acc_s, s_idx, lane);
    }
    __syncthreads();
    for (int j = tid; j < ncoef; j += 128) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += s_acc[w * ncoef_pad + j];
        partial[(size_t)blockIdx.x * ncoef + j] = v;
    }
}

// out[c] = sum_b partial[b][c]; the last launch of the pipeline also resets the task counter
__global__ void __launch_bounds__(256)
sl_sorted_reduce_kernel(const double* __restrict__ partial, int nrows, int ncol, double* __restrict__ out,
                        unsigned int* __restrict__ counter) {
    __shared__ double s_p[32][9];
    const int c = blockIdx.x * 8 + (threadIdx.x & 7), slice = threadIdx.x >> 3;
    double s0 = 0.0, s1 = 0.0;
    if (c < ncol) {
        int b = slice;
        for (; b + 32 < nrows; b += 64) {
            s0 += __ldg(partial + (size_t)b * ncol + c);
            s1 += __ldg(partial + (size_t)(b + 32) * ncol + c);
        }
        if (b < nrows) s0 += __ldg(partial + (size_t)b * ncol + c);
    }
    s_p[slice][threadIdx.x & 7] = s0 + s1;
    __syncthreads();
    if (threadIdx.x < 8 && blockIdx.x * 8 + threadIdx.x < ncol) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += s_p[k][threadIdx.x];
        out[blockIdx.x * 8 + threadIdx.x] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { counter[1] = 0u; counter[2] = 0u; }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int g_bfe_sl_flush_cost = 32;          // option "sl_flush_cost": cost of a run flush in records, for the static split of the deposit kernel
static size_t sl_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SlSortWs {
    int* hist; int* bin_start; int* cursor; SlRec* rec;
    int* binid; int* H;          // stable tile sort: interval per particle, [tiles][nbin] counts
    unsigned int* gprefix;       // cost prefix over the granules of the sorted array (static deposit split)
};

static int64_t sl_tile_rows_for(const bfe_sl* h, int64_t cap) {
    const int64_t a = h->num_sms, b = cap / 8192 + 2;
    return a > b ? a : b;
}

static int sl_sort_workspace(bfe_sl* h, int64_t n, SlSortWs* ws) {
    const int nbin = h->g.numr - 1;
    size_t o_hist = 0;
    size_t o_start = sl_align_up(o_hist + sizeof(int) * nbin, 256);
    size_t o_cur = sl_align_up(o_start + sizeof(int) * (nbin + 1), 256);
    size_t o_rec = sl_align_up(o_cur + sizeof(int) * nbin * BFE_CURSOR_STRIDE, 256);
    if (n > h->sort_cap || !h->sort_ws) {
        if (h->sort_ws) { BFE_CUDA(cudaDeviceSynchronize()); BFE_CUDA(cudaFree(h->sort_ws)); h->sort_ws = nullptr; }
        int64_t cap = n + n / 8 + 1024;
        BFE_CUDA(cudaMalloc(&h->sort_ws, o_rec + (sizeof(SlRec) + sizeof(int)) * (size_t)cap + 256 +
                                             sizeof(int) * (size_t)nbin * (size_t)sl_tile_rows_for(h, cap) + 512 +
                                             sizeof(unsigned int) * (BFE_SL_MAX_GRANULES + 256)));
        BFE_CUDA(cudaMemset(h->sort_ws, 0, o_rec));
        BFE_CUDA(cudaDeviceSynchronize());
        h->sort_cap = cap;
    }
    char* b = (char*)h->sort_ws;
    ws->hist = (int*)(b + o_hist); ws->bin_start = (int*)(b + o_start); ws->cursor = (int*)(b + o_cur);
    ws->rec = (SlRec*)(b + o_rec);
    ws->binid = (int*)(b + sl_align_up(o_rec + sizeof(SlRec) * (size_t)h->sort_cap, 256));
    ws->H = (int*)(b + sl_align_up(o_rec + (sizeof(SlRec) + sizeof(int)) * (size_t)h->sort_cap + 256, 256));
    ws->gprefix = (unsigned int*)(b + sl_align_up(o_rec + (sizeof(SlRec) + sizeof(int)) * (size_t)h->sort_cap + 256 +
                                                  sizeof(int) * (size_t)nbin * (size_t)sl_tile_rows_for(h, h->sort_cap) + 256, 256));
    return BFE_OK;
}

template <int LCAP, int KC>
static int sl_deposit_launch(bfe_sl* h, int64_t n, const SlRec* rec, unsigned int* gprefix, int no_odd, double* expcoef,
                             cudaStream_t stream) {
    constexpr int NW = 4;
    const size_t smem = ((size_t)NW * (2 * (LCAP + 1) * 32) + (size_t)NW * ((LCAP + 1) * (LCAP + 1) + 2) * 33 +
                         (size_t)NW * 128) * sizeof(double) + NW * sizeof(unsigned long long) + 32 * KC * sizeof(int) + 128;
    auto kern = sl_deposit_kernel<LCAP, KC>;
    static bool attr_set = false;                 // one flag per template instance
    if (!attr_set) {
        BFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int ncoef = h->g.nrow * h->g.nmax;
    int64_t nblk = (n + 128 * NW - 1) / (128 * NW);
    int grid = (int)(nblk < (int64_t)h->num_sms * 3 ? nblk : (int64_t)h->num_sms * 3);
    if (grid < 1) grid = 1;
    if (grid > h->max_ctas) grid = h->max_ctas;
    const int use_tma = (h->g.ln % 2 == 0) ? 1 : 0;      // bulk copies need 16-byte aligned rows
    const int static_tasks = (g_bfe_sort_stable && n > 0) ? 1 : 0;
    int gran = 128, ngran = 0;
    if (static_tasks) {
        // granules of 128 records, coarser for very large sets (<= 128 blocks of 1024 granules)
        while ((n + gran - 1) / gran > BFE_SL_MAX_GRANULES) gran *= 2;
        ngran = (int)((n + gran - 1) / gran);
        sl_granule_cost_kernel<<<(ngran + 1023) / 1024, 1024, 0, stream>>>(n, gran, ngran, g_bfe_sl_flush_cost, rec, gprefix,
                                                                          gprefix + BFE_SL_MAX_GRANULES);
        BFE_LAUNCH_CHECK("sl_granule_cost_kernel");
    }
    kern<<<grid, 128, smem, stream>>>(h->g, h->e_node, h->fac, no_odd, n, rec, h->partial, h->counter, use_tma, static_tasks,
                                      gran, ngran, gprefix);
    BFE_LAUNCH_CHECK("sl_deposit_kernel");
    sl_sorted_reduce_kernel<<<(ncoef + 7) / 8, 256, 0, stream>>>(h->partial, grid, ncoef, expcoef, h->counter);
    BFE_LAUNCH_CHECK("sl_sorted_reduce_kernel");
    return BFE_OK;
}

template <int LCAP, int NSPLIT>
static int sl_deposit_lane_launch(bfe_sl* h, int64_t n, const SlRec* rec, int no_odd, double* expcoef, cudaStream_t stream) {
    constexpr int NW = 4, RS = 33;
    constexpr int NR0 = (NSPLIT == 1) ? SlLaneRows<LCAP, 0, LCAP>::count() : SlLaneRows<LCAP, 0, 2>::count();
    const int ncoef = h->g.nrow * h->g.nmax, ncoef_pad = (ncoef + 31) & ~31;
    const size_t smem = ((size_t)NW * 2 * NR0 * RS + (size_t)NW * 128 + (size_t)NW * ncoef_pad) * sizeof(double) +
                        (size_t)ncoef_pad * sizeof(int) + 64;
    auto kern = sl_deposit_lane_kernel<LCAP, NSPLIT>;
    static size_t attr_smem = 0;                   // one per template instance
    if (smem > attr_smem) {
        BFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    int64_t nblk = (n + 1024 * NW - 1) / (1024 * NW) * NSPLIT;
    int grid = (int)(nblk < (int64_t)h->num_sms * 2 ? nblk : (int64_t)h->num_sms * 2);
    if (grid < 1) grid = 1;
    if (grid > h->max_ctas) grid = h->max_ctas;
    kern<<<grid, 128, smem, stream>>>(h->g, h->e_node, h->fac, no_odd, n, rec, h->partial, h->counter);
    BFE_LAUNCH_CHECK("sl_deposit_lane_kernel");
    sl_sorted_reduce_kernel<<<(ncoef + 7) / 8, 256, 0, stream>>>(h->partial, grid, ncoef, expcoef, h->counter);
    BFE_LAUNCH_CHECK("sl_sorted_reduce_kernel");
    return BFE_OK;
}

bool bfe_sl_sorted_supported(const bfe_sl* h) {
    const int ncoef = h->g.nrow * h->g.nmax;
    if (h->g.lmax <= 4 && ncoef <= 32 * 15) return true;
    if (h->g.lmax <= 6 && ncoef <= 32 * 28) return true;
    return false;
}

int bfe_sl_accumulate_sorted(bfe_sl* h, int64_t n, const double* x, const double* y, const double* z,
                             const double* mass, int no_odd, double* expcoef, cudaStream_t stream) {
    if (!bfe_sl_sorted_supported(h)) return BFE_ERR_UNSUPPORTED;
    if (n >= (int64_t)1 << 31) return BFE_ERR_UNSUPPORTED;
    SlSortWs ws;
    int rc = sl_sort_workspace(h, n, &ws);
    if (rc != BFE_OK) return rc;
    const int nbin = h->g.numr - 1;
    bool sorted_done = false;
    if (g_bfe_sort_stable && n > 0 && g_bfe_sl_deposit_mode != 2) {
        int64_t U = (n + (int64_t)h->num_sms * 1024 - 1) / ((int64_t)h->num_sms * 1024);
        U = U < 1 ? 1 : (U > 8 ? 8 : U);
        const int tile = (int)(U * 1024);
        const int ntile = (int)((n + tile - 1) / tile);
        const int per = (nbin + 1023) / 1024;
        const size_t ss_h = sizeof(int) * (size_t)per * 1024;
        const size_t ss_s = ss_h + sizeof(unsigned short) * (size_t)tile + 16;
        if (ss_s <= 200 * 1024 && (int64_t)ntile <= sl_tile_rows_for(h, h->sort_cap) && (nbin + 63) / 64 <= 1024) {
            if (ss_h > 48 * 1024) BFE_CUDA(cudaFuncSetAttribute(sl_tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss_h));
            if (ss_s > 48 * 1024) BFE_CUDA(cudaFuncSetAttribute(sl_tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss_s));
            sl_tile_hist_kernel<<<ntile, 1024, ss_h, stream>>>(h->g, h->xi, nbin, n, tile, x, y, z, ws.H, ws.binid);
            BFE_LAUNCH_CHECK("sl_tile_hist_kernel");
            rc = bfe_tile_colscan(nbin, ntile, ws.H, ws.hist, ws.bin_start, h->counter, stream);
            if (rc != BFE_OK) return rc;
            sl_tile_scatter_kernel<<<ntile, 1024, ss_s, stream>>>(h->g, h->xi, h->p0, nbin, n, tile, x, y, z, mass,
                                                                 (const int*)ws.bin_start, (const int*)ws.H, ws.rec, (const int*)ws.binid);
            BFE_LAUNCH_CHECK("sl_tile_scatter_kernel");
            sorted_done = true;
        }
    }
    int grid = (int)((n + 1023) / 1024);
    if (grid > h->num_sms * 2) grid = h->num_sms * 2;
    if (grid < 1) grid = 1;
    if (!sorted_done) {
        const int per = (nbin + 1023) / 1024;
        const size_t ss = sizeof(int) * (size_t)per * 1024;
        if (ss > 200 * 1024) return BFE_ERR_UNSUPPORTED;
        if (ss > 48 * 1024)
            BFE_CUDA(cudaFuncSetAttribute(sl_bin_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss));
        sl_bin_hist_kernel<<<grid, 1024, ss, stream>>>(h->g, h->xi, nbin, n, x, y, z, ws.hist, ws.bin_start, ws.cursor,
                                                      h->counter);
    }
    BFE_LAUNCH_CHECK("sl_bin_hist_kernel");
    int g2 = (int)((n + 511) / 512);
    if (g2 > h->num_sms * 8) g2 = h->num_sms * 8;
    if (g2 < 1) g2 = 1;
    if (!sorted_done) {
        sl_bin_scatter_kernel<<<g2, 256, 0, stream>>>(h->g, h->xi, h->p0, n, x, y, z, mass, ws.bin_start, ws.cursor, ws.rec);
        BFE_LAUNCH_CHECK("sl_bin_scatter_kernel");
    }
    if (g_bfe_sl_deposit_mode == 2) {             // register formulation (option sl_deposit_mode = 2); default: slab kernel
        if (h->g.lmax <= 4) return sl_deposit_lane_launch<4, 1>(h, n, ws.rec, no_odd, expcoef, stream);
        return sl_deposit_lane_launch<6, 2>(h, n, ws.rec, no_odd, expcoef, stream);
    }
    if (h->g.lmax <= 4 && h->g.nrow * h->g.nmax <= 32 * 15)
        return sl_deposit_launch<4, 15>(h, n, ws.rec, ws.gprefix, no_odd, expcoef, stream);
    return sl_deposit_launch<6, 28>(h, n, ws.rec, ws.gprefix, no_odd, expcoef, stream);
}
