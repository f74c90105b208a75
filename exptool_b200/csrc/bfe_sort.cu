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
// (integer keys, integer cursors), the 52 sums are formed per run of equal cells with
// per-thread register accumulators and a shared-memory combine, and the table rows are
// read once per RUN instead of once per particle.  No floating-point atomics anywhere;
// the coefficient partials are combined exactly as in the direct kernel.
//
//   eof_cell_hist_kernel     : per-particle cell id -> histogram (shared-memory int counters);
//                              the last CTA to finish scans it (cell_start, cursors)
//   eof_cell_scatter_kernel  : 64-byte record {c00,c10,c01,c11,cos phi,sin phi,m|R,cell:perm}
//                              written at its sorted position
//   eof_deposit_kernel       : runs -> S -> coefficient partials -> last-CTA reduce
//   eof_force_sorted_kernel  : field evaluation in sorted order (warp-uniform table rows),
//                              outputs scattered back to the caller's particle order
#include "bfe_device.cuh"
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
                     int* __restrict__ cursor, unsigned int* __restrict__ counter) {
    extern __shared__ int s_hist[];
    __shared__ int s_wsum[32];
    __shared__ bool s_last;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) s_hist[c] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        double r = sqrt(px * px + py * py + 1.e-10);
        EofBin b = bfe_eof_bin(g, r, pz);
        atomicAdd(&s_hist[bfe_cell_of(g, b)], 1);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
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
        bfe_block_scan_cells(ncell, hist, cell_start, cursor, s_hist, s_wsum);
        if (threadIdx.x == 0) *counter = 0u;
    }
}

__global__ void __launch_bounds__(256)
eof_cell_scatter_kernel(EofGeom g, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                        const double* __restrict__ z, const double* __restrict__ mass,
                        int* __restrict__ cursor, EofRec* __restrict__ rec, int* __restrict__ inv,
                        double* __restrict__ r_orig) {
    constexpr int U = 2;       // particles per thread per pass: independent slot claims overlap their latency
    for (int64_t base = (int64_t)blockIdx.x * (256 * U); base < n; base += (int64_t)gridDim.x * (256 * U)) {
        EofBin b[U];
        double c1[U], s1[U], r[U], aux[U];
        int cell[U], pos[U];
        int64_t idx[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            idx[u] = base + u * 256 + threadIdx.x;
            const bool on = idx[u] < n;
            double px = on ? __ldg(x + idx[u]) : 1.0, py = on ? __ldg(y + idx[u]) : 0.0, pz = on ? __ldg(z + idx[u]) : 0.0;
            aux[u] = (on && mass) ? __ldg(mass + idx[u]) : 0.0;
            r[u] = sqrt(px * px + py * py + 1.e-10);              // eof.py:531 / 1070
            b[u] = bfe_eof_bin(g, r[u], pz);
            cell[u] = bfe_cell_of(g, b[u]);
            bfe_cossin_phi(px, py, c1[u], s1[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            pos[u] = (idx[u] < n) ? atomicAdd(&cursor[cell[u]], 1) : 0;   // integer slot claim, not a data reduction
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (idx[u] < n) {
                unsigned long long cp = ((unsigned long long)(unsigned int)cell[u] << 32) |
                                        (unsigned long long)(unsigned int)idx[u];
                double2* dst = reinterpret_cast<double2*>(rec + pos[u]);
                dst[0] = make_double2(b[u].c00, b[u].c10);
                dst[1] = make_double2(b[u].c01, b[u].c11);
                dst[2] = make_double2(c1[u], s1[u]);
                dst[3] = make_double2(aux[u], __longlong_as_double((long long)cp));
                inv[idx[u]] = pos[u];             // original index -> sorted slot (coalesced)
                r_orig[idx[u]] = r[u];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// deposit + contract over sorted records.
//
// A WARP owns a task of TASK consecutive sorted records and walks it 32 records at a time:
//   1. lanes = records: the 64-B record (prefetched one batch ahead) is expanded to 4 mass-weighted
//      corner weights and 2 mmax+1 cos/sin(m phi) values in the warp's shared-memory slab;
//   2. the per-run sums  S[k][ch] = sum_p w_k(p) trig_ch(p)  are a (4 x L)(L x ntrig) product:
//      done on the FP64 tensor cores with mma.sync.m8n8k4 (A = W^T, rows 4..7 zero; B = trig,
//      two n-tiles of 8 channels; 4 records per k-step).  The direct LDS+DFMA form needs 8
//      shared-memory wavefronts per record and was bound on shared-memory bandwidth (ncu:
//      profiles/); the fragment loads need ~1.3.  Records outside the current run are masked to
//      zero in the A fragment, so a run boundary costs at most one extra k-step;
//   3. when the cell id changes (warp-uniform) the finished run is flushed: every lane adds
//      sum_k T[node_k][j] S[k][ch(j)] to the register accumulators of its channels j = lane + 32 c.
// Batches that are mostly single-record runs (sparse outskirts) are deferred and done at the end
// by the whole CTA in the direct formulation (thread per channel), which keeps all 8 warps' load
// pipelines busy instead of serialising ~30 dependent flushes in one warp.
// ---------------------------------------------------------------------------
// Optional per-warp cycle counters (build with BFE_NVCC_FLAGS=-DBFE_PROFILE_DEPOSIT; profiles/deposit_cycles.py)
#ifdef BFE_PROFILE_DEPOSIT
__device__ long long* g_dbg = nullptr;
extern "C" int bfe_debug_set(long long* p) { return (int)cudaMemcpyToSymbol(g_dbg, &p, sizeof(p)); }
#define DBG_DECL long long dbg_t0 = clock64(), dbg_tk = 0, dbg_flush = 0, dbg_mma = 0, dbg_expand = 0, dbg_nflush = 0, dbg_nks = 0, dbg_wait = 0
#define DBG_TICK() (dbg_tk = clock64())
#define DBG_ADD(var) do { long long t_ = clock64(); var += t_ - dbg_tk; dbg_tk = t_; } while (0)
#else
#define DBG_DECL
#define DBG_TICK()
#define DBG_ADD(var)
#endif

struct DepositSmem {
    static constexpr int NW = 8;                  // warps per CTA
    static constexpr int RS = 36;                 // slab row stride (doubles): conflict-free fragment loads
    static constexpr int TROW = 256;              // max padded channels per node row
    static constexpr int MAXDEFER = 160;
};

template <int MCAP, int KCH>
__global__ void __launch_bounds__(256, 2)
eof_deposit_kernel(EofGeom g, const double* __restrict__ t_acc, int nch, int nch_pad, int64_t n,
                   const EofRec* __restrict__ rec, double* __restrict__ partial, unsigned int* __restrict__ counter,
                   double* __restrict__ cos_out, double* __restrict__ sin_out) {
    constexpr int TASK = 128;
    constexpr int SPARSE_RUNS = 20;               // batches with more run starts than this are deferred
    constexpr int NTRIG = 2 * MCAP + 1;
    constexpr int NV = 4 + NTRIG;                 // values per record in the slab
    constexpr int NW = DepositSmem::NW, RS = DepositSmem::RS, TROW = DepositSmem::TROW;
    constexpr int MAXDEFER = DepositSmem::MAXDEFER;
    constexpr int SLAB = NV * RS;
    static_assert(SLAB >= KCH * 32, "slab reused for the CTA reduce");
    static_assert(NTRIG <= 16, "two n-tiles of 8 channels");
    // dynamic shared memory carve-up (about 107 kB; two CTAs per SM)
    extern __shared__ __align__(128) unsigned char s_raw[];
    double* s_tbuf = reinterpret_cast<double*>(s_raw);                    // [NW][4][TROW] table rows of the open run
    double* s_slab = s_tbuf + NW * 4 * TROW;                              // [NW][SLAB]
    double* s_S = s_slab + NW * SLAB;                                     // [NW][64]
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_S + NW * 64);   // [NW] mbarriers
    int* s_defer = reinterpret_cast<int*>(s_bar + NW);                    // [MAXDEFER]
    int* s_dcell = s_defer + MAXDEFER;                                    // [32]
    __shared__ int s_ndefer;
    __shared__ bool s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ntrig = 2 * g.mmax + 1;
    const int ncos = (g.mmax + 1) * g.norder;
    const int fg = lane >> 2, fj = lane & 3;      // mma fragment coordinates: group, thread-in-group
    const bool a_on = fg < 4;                     // A rows 4..7 are zero
    const bool b1_on = (8 + fg) < ntrig;          // second n-tile holds channels 8..ntrig-1
    // channels owned by this lane: j = lane + 32 c
    int trig_idx[KCH];
#pragma unroll
    for (int c = 0; c < KCH; ++c) {
        int j = lane + 32 * c;
        trig_idx[c] = (j < nch) ? ((j < ncos) ? j / g.norder : g.mmax + 1 + (j - ncos) / g.norder) : 0;
    }
    double acc[KCH];
#pragma unroll
    for (int c = 0; c < KCH; ++c) acc[c] = 0.0;
    const int rowstep = nch_pad, colstep = g.ny1 * nch_pad;
    const unsigned int rowbytes = 2u * (unsigned int)nch_pad * 8u;        // nodes (ix,iy),(ix,iy+1) are contiguous

    double* val = s_slab + warp * SLAB;           // val[v * RS + p]
    double* S = s_S + warp * 64;
    double* tb = s_tbuf + warp * 4 * TROW;        // [T00 | T01] [T10 | T11], each nch_pad wide
    const unsigned int bar = bfe_smem_u32(s_bar + warp);
    const unsigned int tb_u32 = bfe_smem_u32(tb);
    unsigned int bar_parity = 0;
    if (tid == 0) s_ndefer = 0;
    if (lane == 0) bfe_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // warps pull tasks from a global counter (counter[1]); the last CTA resets it
    unsigned int* task_counter = counter + 1;
    const int64_t ntasks = (n + TASK - 1) / TASK;
    DBG_DECL;
    for (;;) {
        int64_t task = 0;
        if (lane == 0) task = (int64_t)atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntasks) break;
        const int64_t t0 = task * TASK;
        const int tcnt = (int)((n - t0) < TASK ? (n - t0) : TASK);
        double c00 = 0.0, c01 = 0.0, c10 = 0.0, c11 = 0.0;   // C fragments: tile 0 (ch 0..7), tile 1 (ch 8..15)
        int cur = -1;                               // cell of the run being summed (-1: none)
        double2 ra, rb, rc, rd;
        {
            const bool on = lane < ((tcnt < 32) ? tcnt : 32);
            const double2* src = reinterpret_cast<const double2*>(rec + t0 + (on ? lane : 0));
            ra = __ldg(src); rb = __ldg(src + 1); rc = __ldg(src + 2); rd = __ldg(src + 3);
        }
        for (int b0 = 0; b0 < tcnt; b0 += 32) {
            const int bcnt = (tcnt - b0) < 32 ? (tcnt - b0) : 32;
            __syncwarp();
            DBG_TICK();
            int mycell = -3;
            {
                // lanes >= bcnt park zeros so that k-steps may read the whole 32-record slab
                const bool on = lane < bcnt;
                const double m = on ? rd.x : 0.0;
#ifdef BFE_PROFILE_DEPOSIT
                if (__double_as_longlong(m) == 0x7ff8dead00000000ll) dbg_wait = 1;   // force the load to complete
                DBG_ADD(dbg_wait);
#endif
                val[0 * RS + lane] = ra.x * m; val[1 * RS + lane] = ra.y * m;
                val[2 * RS + lane] = rb.x * m; val[3 * RS + lane] = rb.y * m;
                if (on) mycell = (int)((unsigned long long)__double_as_longlong(rd.y) >> 32);
                double cm = 1.0, sm = 0.0;
                val[4 * RS + lane] = 1.0;
#pragma unroll
                for (int mm = 1; mm <= MCAP; ++mm) {
                    if (mm <= g.mmax) {
                        double cn = cm * rc.x - sm * rc.y, sn = sm * rc.x + cm * rc.y;
                        cm = cn; sm = sn;
                        val[(4 + mm) * RS + lane] = cm;
                        val[(4 + g.mmax + mm) * RS + lane] = sm;
                    }
                }
            }
            if (b0 + 32 < tcnt) {
                const bool on = (b0 + 32 + lane) < tcnt;
                const double2* src = reinterpret_cast<const double2*>(rec + t0 + b0 + 32 + (on ? lane : 0));
                ra = __ldg(src); rb = __ldg(src + 1); rc = __ldg(src + 2); rd = __ldg(src + 3);
            }
            // run boundaries inside this batch (bit p set: record p starts a new run)
            int prevcell = __shfl_up_sync(0xffffffffu, mycell, 1);
            if (lane == 0) prevcell = cur;
            unsigned int bmask = __ballot_sync(0xffffffffu, (lane < bcnt) && (mycell != prevcell));
            __syncwarp();
            DBG_ADD(dbg_expand);

            // open a run: one lane issues two TMA bulk copies that bring the run's four table rows into the
            // warp's buffer while the tensor cores sum the run; the flush then reads them from shared memory
#define BFE_OPEN_RUN(cellv)                                                                               \
            do {                                                                                          \
                cur = (cellv);                                                                            \
                if (lane == 0) {                                                                          \
                    const int ox_ = cur / g.numy, oy_ = cur - ox_ * g.numy;                               \
                    const double* src_ = t_acc + (size_t)(ox_ * g.ny1 + oy_) * nch_pad;                   \
                    bfe_mbar_expect_tx(bar, 2u * rowbytes);                                               \
                    bfe_bulk_g2s(tb_u32, src_, rowbytes, bar);                                            \
                    bfe_bulk_g2s(tb_u32 + 2u * TROW * 8u, src_ + colstep, rowbytes, bar);                 \
                }                                                                                         \
            } while (0)

#define BFE_FLUSH_RUN()                                                                                   \
            do {                                                                                          \
                __syncwarp();                                                                             \
                if (a_on) {                                                                               \
                    S[fg * 16 + 2 * fj] = c00; S[fg * 16 + 2 * fj + 1] = c01;                             \
                    S[fg * 16 + 8 + 2 * fj] = c10; S[fg * 16 + 8 + 2 * fj + 1] = c11;                     \
                }                                                                                         \
                bfe_mbar_wait(bar, bar_parity);                                                           \
                bar_parity ^= 1u;                                                                         \
                __syncwarp();                                                                             \
                _Pragma("unroll")                                                                         \
                for (int c = 0; c < KCH; ++c) {                                                           \
                    const int j_ = lane + 32 * c;                                                         \
                    if (j_ < nch) {                                                                       \
                        const int ti_ = trig_idx[c];                                                      \
                        acc[c] += tb[j_] * S[ti_] + tb[2 * TROW + j_] * S[16 + ti_] +                     \
                                  tb[rowstep + j_] * S[32 + ti_] + tb[2 * TROW + rowstep + j_] * S[48 + ti_]; \
                    }                                                                                     \
                }                                                                                         \
                __syncwarp();                                                                             \
                c00 = 0.0; c01 = 0.0; c10 = 0.0; c11 = 0.0;                                               \
            } while (0)

            if (__popc(bmask) > SPARSE_RUNS) {
                // ---- sparse batch: close the open run and hand the batch to the CTA-wide direct pass
                int slot = 0;
                if (lane == 0) slot = atomicAdd(&s_ndefer, 1);
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (slot < MAXDEFER) {
                    if (cur >= 0) { BFE_FLUSH_RUN(); cur = -1; }
                    if (lane == 0) s_defer[slot] = (int)(t0 + b0);       // n < 2^31
                    continue;
                }
                // queue full (pathological input): fall through and do it here, run by run
            }

            int p = 0;
            while (p < bcnt) {
                if ((bmask >> p) & 1u) {
                    DBG_TICK();
                    if (cur >= 0) {
                        BFE_FLUSH_RUN();
#ifdef BFE_PROFILE_DEPOSIT
                        dbg_nflush++;
                        if (__double_as_longlong(acc[0]) == 0x7ff8dead00000000ll) dbg_wait = 1;
#endif
                    }
                    BFE_OPEN_RUN(__shfl_sync(0xffffffffu, mycell, p));
                    DBG_ADD(dbg_flush);
                }
                DBG_TICK();
                // end of this run within the batch: next boundary after p, or bcnt
                const unsigned int later = (p < 31) ? (bmask >> (p + 1)) : 0u;
                const int qend = later ? (p + __ffs(later)) : bcnt;
                // k-steps of 4 records covering [p, qend); records outside the run are masked in A
                for (int ks = p >> 2; ks * 4 < qend; ++ks) {
                    const int r = ks * 4 + fj;
                    const bool in = (r >= p) && (r < qend);
                    const double a = (a_on && in) ? val[fg * RS + r] : 0.0;
                    const double bt0 = val[(4 + fg) * RS + r];
                    const double bt1 = b1_on ? val[(12 + fg) * RS + r] : 0.0;
                    bfe_dmma_m8n8k4(c00, c01, a, bt0);
                    bfe_dmma_m8n8k4(c10, c11, a, bt1);
#ifdef BFE_PROFILE_DEPOSIT
                    dbg_nks++;
#endif
                }
#ifdef BFE_PROFILE_DEPOSIT
                if (__double_as_longlong(c00) == 0x7ff8dead00000000ll) dbg_wait = 1;
#endif
                DBG_ADD(dbg_mma);
                p = qend;
            }
        }
        if (cur >= 0) BFE_FLUSH_RUN();
#undef BFE_FLUSH_RUN
#undef BFE_OPEN_RUN
    }

#ifdef BFE_PROFILE_DEPOSIT
    long long dbg_main = clock64() - dbg_t0;
#endif
    // ---- deferred sparse batches: direct formulation, thread per channel, whole CTA
    __syncthreads();
    double accd = 0.0;
    {
        const int nd = s_ndefer < MAXDEFER ? s_ndefer : MAXDEFER;
        int my_ti = 0;
        if (tid < nch) my_ti = (tid < ncos) ? tid / g.norder : g.mmax + 1 + (tid - ncos) / g.norder;
        double* v0 = s_slab;
        for (int d = 0; d < nd; ++d) {
            const int64_t r0 = s_defer[d];
            const int bcnt = (int)((n - r0) < 32 ? (n - r0) : 32);
            const int64_t tend = (r0 / TASK + 1) * (int64_t)TASK;           // batches never cross a task
            const int cnt = (int)((tend - r0) < bcnt ? (tend - r0) : bcnt);
            __syncthreads();
            if (tid < 32) {
                const bool on = tid < cnt;
                const double2* src = reinterpret_cast<const double2*>(rec + r0 + (on ? tid : 0));
                double2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), dd = __ldg(src + 3);
                const double m = on ? dd.x : 0.0;
                v0[0 * RS + tid] = a.x * m; v0[1 * RS + tid] = a.y * m; v0[2 * RS + tid] = b.x * m; v0[3 * RS + tid] = b.y * m;
                const int cell = (int)((unsigned long long)__double_as_longlong(dd.y) >> 32);
                const int ix = cell / g.numy, iy = cell - ix * g.numy;
                s_dcell[tid] = on ? (ix * g.ny1 + iy) : 0;
                double cm = 1.0, sm = 0.0;
                v0[4 * RS + tid] = 1.0;
#pragma unroll
                for (int mm = 1; mm <= MCAP; ++mm) {
                    if (mm <= g.mmax) {
                        double cn = cm * c.x - sm * c.y, sn = sm * c.x + cm * c.y;
                        cm = cn; sm = sn;
                        v0[(4 + mm) * RS + tid] = cm;
                        v0[(4 + g.mmax + mm) * RS + tid] = sm;
                    }
                }
            }
            __syncthreads();
            if (tid < nch) {
                const double* tcol = t_acc + tid;
                const double* trow = v0 + (4 + my_ti) * RS;
#pragma unroll 4
                for (int p = 0; p < 32; ++p) {
                    const double* base = tcol + (size_t)s_dcell[p] * nch_pad;
                    double v = __ldg(base) * v0[0 * RS + p] + __ldg(base + colstep) * v0[1 * RS + p] +
                               __ldg(base + rowstep) * v0[2 * RS + p] + __ldg(base + colstep + rowstep) * v0[3 * RS + p];
                    accd = fma(trow[p], v, accd);
                }
            }
        }
    }

#ifdef BFE_PROFILE_DEPOSIT
    if (lane == 0 && g_dbg) {
        long long* d = g_dbg + (blockIdx.x * NW + warp) * 8;
        d[0] = dbg_main; d[1] = clock64() - dbg_t0; d[2] = dbg_wait; d[3] = dbg_expand; d[4] = dbg_mma;
        d[5] = dbg_flush; d[6] = dbg_nflush; d[7] = dbg_nks;
    }
#endif
    // ---- CTA reduce over the 8 warps, then per-CTA partial and last-CTA final reduce
    __syncthreads();
#pragma unroll
    for (int c = 0; c < KCH; ++c) s_slab[warp * SLAB + c * 32 + lane] = acc[c];
    __syncthreads();
    for (int j = tid; j < nch_pad; j += 256) {
        double v = (j == tid) ? accd : 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += s_slab[w * SLAB + j];
        partial[(size_t)blockIdx.x * nch_pad + j] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int j = tid; j < nch; j += 256) {
            double s = bfe_column_sum(partial, (int)gridDim.x, nch_pad, j) * BFE_FOURPI_NEG;
            if (j < ncos) cos_out[j] = s;
            else          sin_out[j - ncos + g.norder] = s;
        }
        for (int k = tid; k < g.norder; k += 256) sin_out[k] = 0.0;
        if (tid == 0) { counter[0] = 0u; counter[1] = 0u; }
    }
}

static size_t deposit_smem_bytes() {
    const int NV = 4 + 13;
    size_t doubles = (size_t)DepositSmem::NW * 4 * DepositSmem::TROW + (size_t)DepositSmem::NW * NV * DepositSmem::RS +
                     (size_t)DepositSmem::NW * 64;
    return doubles * 8 + DepositSmem::NW * 8 + (DepositSmem::MAXDEFER + 32) * 4 + 128;
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

__global__ void __launch_bounds__(256)
eof_force_gather_kernel(int64_t n, const int* __restrict__ inv, const double* __restrict__ r_orig,
                        const double2* __restrict__ tmp, double* __restrict__ p0, double* __restrict__ p,
                        double* __restrict__ fr, double* __restrict__ fp, double* __restrict__ fz,
                        double* __restrict__ R) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2* src = tmp + 3 * (int64_t)__ldg(inv + i);
        const double2 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2);
        p0[i] = a.x; p[i] = a.y; fr[i] = b.x; fp[i] = b.y; fz[i] = c.x; R[i] = __ldg(r_orig + i);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SortWs {
    int* hist; int* cell_start; int* cursor; EofRec* rec; double* r_orig; double2* tmp; int* inv;
};

static int sort_workspace(bfe_eof* h, int64_t n, SortWs* ws) {
    const int ncell = h->g.numx * h->g.numy;
    size_t o_hist = 0;
    size_t o_start = align_up(o_hist + sizeof(int) * ncell, 256);
    size_t o_cur = align_up(o_start + sizeof(int) * (ncell + 1), 256);
    size_t o_rec = align_up(o_cur + sizeof(int) * ncell, 256);
    if (n > h->sort_cap || !h->sort_ws) {
        if (h->sort_ws) { BFE_CUDA(cudaDeviceSynchronize()); BFE_CUDA(cudaFree(h->sort_ws)); h->sort_ws = nullptr; }
        int64_t cap = (n + n / 8 + 1024 + 15) / 16 * 16;      // multiple of 16: keeps every sub-array 16-B aligned
        // per particle: 64-B record, R (8 B), 48-B force slot, inverse permutation (4 B)
        BFE_CUDA(cudaMalloc(&h->sort_ws, o_rec + (sizeof(EofRec) + sizeof(double) + 48 + sizeof(int)) * (size_t)cap));
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
    int grid = (int)((n + 1023) / 1024);
    if (grid > h->num_sms * 2) grid = h->num_sms * 2;
    if (grid < 1) grid = 1;
    {
        const int per = (ncell + 1023) / 1024;
        const size_t ss = sizeof(int) * (size_t)per * 1024;
        if (ss > 200 * 1024) return BFE_ERR_UNSUPPORTED;
        if (ss > 48 * 1024)
            BFE_CUDA(cudaFuncSetAttribute(eof_cell_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss));
        const int kt = bfe_kt_begin("eof_cell_hist_kernel", stream);
        eof_cell_hist_kernel<<<grid, 1024, ss, stream>>>(h->g, ncell, n, x, y, z, ws.hist, ws.cell_start, ws.cursor,
                                                        h->counter);
        bfe_kt_end(kt, stream);
    }
    BFE_LAUNCH_CHECK("eof_cell_hist_kernel");
    int g2 = (int)((n + 511) / 512);
    if (g2 > h->num_sms * 8) g2 = h->num_sms * 8;
    if (g2 < 1) g2 = 1;
    const int kt2 = bfe_kt_begin("eof_cell_scatter_kernel", stream);
    eof_cell_scatter_kernel<<<g2, 256, 0, stream>>>(h->g, n, x, y, z, mass, ws.cursor, ws.rec, ws.inv, ws.r_orig);
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
    int64_t nblk = (n + 128 * 8 - 1) / (128 * 8);          // 8 warp tasks of 128 records per CTA pass
    int grid = (int)(nblk < (int64_t)h->num_sms * 2 ? nblk : (int64_t)h->num_sms * 2);
    if (grid < 1) grid = 1;
    {
        static bool attr_set = false;
        const size_t smem = deposit_smem_bytes();
        if (!attr_set) {
            BFE_CUDA(cudaFuncSetAttribute(eof_deposit_kernel<6, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attr_set = true;
        }
        const int kt = bfe_kt_begin("eof_deposit_kernel", stream);
        eof_deposit_kernel<6, 8><<<grid, 256, smem, stream>>>(h->g, h->t_acc, h->nch, h->nch_pad, n, ws.rec, h->partial,
                                                             h->counter, cos_out, sin_out);
        bfe_kt_end(kt, stream);
    }
    BFE_LAUNCH_CHECK("eof_deposit_kernel");
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
    int64_t need = (n + 127) / 128, cap = (int64_t)h->num_sms * 16;
    int grid = (int)(need < cap ? need : cap);
    const int kt = bfe_kt_begin("eof_force_sorted_kernel", stream);
    eof_force_sorted_kernel<6><<<grid, 128, 0, stream>>>(h->g, h->g_con, h->gstride, n, ws.rec, ws.tmp);
    bfe_kt_end(kt, stream);
    BFE_LAUNCH_CHECK("eof_force_sorted_kernel");
    int64_t need2 = (n + 255) / 256, cap2 = (int64_t)h->num_sms * 8;
    const int kt3 = bfe_kt_begin("eof_force_gather_kernel", stream);
    eof_force_gather_kernel<<<(int)(need2 < cap2 ? need2 : cap2), 256, 0, stream>>>(n, ws.inv, ws.r_orig, ws.tmp, p0, p,
                                                                                  fr, fp, fz, R);
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
