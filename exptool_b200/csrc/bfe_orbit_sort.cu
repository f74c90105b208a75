// bfe_orbit_sort.cu -- per-point field evaluation and leapfrog integration with the warps kept TABLE-COHERENT.
//
// The per-point field kernels (bfe_field.cu) are bound by the L1 tag stage: one cycle per distinct 128-byte line per load
// instruction, 84 divergent 256-bit loads per point (ncu l1tex 92 %, FP64 pipe 20 %, profiles/r01_ncu_full_blk_kernels.csv).
// When the 32 lanes of a warp sit in the same EOF (R, z) cell AND the same SL radial interval they read the same two blocks
// (G4[cell], A3[j]): every load instruction touches ONE line, the evaluation becomes FP64-bound.  This file orders the
// points / orbits by a composite key and evaluates in that order WITHOUT moving the data:
//
//   key  = (EOF cell << SUB) | (SL interval & (2^SUB - 1)),  SUB = 4           [FP32 arithmetic: an ordering key only]
//          -- same cell => same G4 block; inside a cell the SL intervals a point can have span a range that depends on the
//          cell (a few for disc-plane cells, ~100 for cells high above the plane): the low SUB bits group equal intervals
//          exactly when the span is <= 16 and leave span/16 distinct intervals per bin otherwise.
//   sort = ONE counting pass over nkeys = ncell * 16 bins (131072 for the 128 x 64 grid), permutation only:
//            pt_key_kernel / orbit_pack_key_kernel : key, rank-in-bin by one integer atomicAdd WITH return  -> (key, rank)
//            key_scan_kernel                       : exclusive scan of the histogram (1024-bin blocks + block prefixes by the
//                                                    last CTA), histogram cleared for the next use
//            perm_scatter_kernel                   : perm[bprefix[key>>10] + start[key] + rank] = i   (no atomics)
//          The order inside a bin follows the atomic's arrival order; the VALUES do not depend on it (each point is evaluated
//          on its own), so results are bit-identical to the caller-order kernels whatever the order.
//   eval = field_rec_kernel / leapfrog_perm_kernel: 128-point tiles of the sorted order are handed out through an atomic
//          ticket (tiles in sparse regions are latency-bound and take longer: with a static split the SMs idled 29 % of the
//          kernel, ncu sm active vs elapsed cycles).  Points: 32-byte records {x,y,z,index} in sorted order in, 64-byte result
//          slots in sorted order out (full sectors, coalesced), field_gather_kernel returns them to the caller's order.
//          Orbits: the 96-byte record (three full sectors) stays in the caller's order and is gathered through the
//          permutation; the key of the NEXT evaluation point (the following step's drift) is produced by the same kernel.
//   What bounds the evaluation once the warps are coherent (ncu, profiles/r02_*): the L1 data pipe -- every 256-bit load
//   is 8 wavefronts (1 KiB into the register file) whatever the coherence, 84 table loads per point -- and then FP64 issue
//   (1700 instructions per point).  The factorial factors therefore travel in the kernel parameter block (SlFacP), not
//   through 28 more loads.  Tried and dropped (profiles/r02_summary.md): two sorted points per lane sharing the loads (halves
//   the L1 traffic, but 224 registers leave two warps per scheduler: 216 -> 278 us per 1e6 disc points) and an L1 prefetch
//   (CCTL.E.PF1) of the next tile's blocks (no gain: the misses are already covered by the four-tile tickets).
//
// Points are processed in chunks (option "field_sort_chunk") so that the chunk's records, slots and keys (100 B per point, a
// workspace that is re-used by every chunk) stay L2-resident: HBM sees the compulsory 24 B in + 64 B out per point.
// Orbits are re-keyed every K steps (option "orbit_resort").  The per-point arithmetic is bfe_field_cart_blk, the same
// function the caller-order kernels use, so results are bit-identical to theirs.
#include "bfe_sortcore.cuh"

#define BFE_KEY_SUBBITS 4
#define BFE_SCAN_BLOCK 1024
#ifndef BFE_PERM_MINB
#define BFE_PERM_MINB 3          // resident CTAs per SM the evaluation kernels are compiled for (4: 128 registers, 3: 168, no spills;
                                 // a quarter of the register file stays free for the support kernels of the other stream)
#endif

__device__ __forceinline__ int bfe_sl_interval_fast(const SlGeom& g, float r) {
    float x;
    if (g.cmap == 1) { const float q = r * (float)g.inv_scale; x = (q - 1.0f) / (q + 1.0f); }
    else if (g.cmap == 2) x = logf(r);
    else x = r;
    int i = (int)((x - (float)g.xi0) * (float)g.inv_dxi);      // NaN -> 0
    if (i < 0) i = 0;
    if (i > g.numr - 2) i = g.numr - 2;
    return i;
}

// The ordering key.  Bit scheme (kc == nullptr): (cell << subbits) | (interval & mask).  Per-cell spans (kc != nullptr, the default):
// kc[cell] = {first key of the cell, first radial interval the cell can reach, number of intervals it spans, -}, so a cell of the
// disc plane (a few intervals) takes a few keys and a cell of the outer halo (hundreds) as many as it needs:
// key = koff + (interval - jmin), clamped; spans above 256 (the open cells at the table's edges) alias modulo 256.  Built once per
// (EOF table, SL table) pair by key_cell_span_kernel / key_cell_scan_kernel.  FP32 arithmetic, an ordering hint only.
// With kc the argument `subbits` is the SHIFT of the slot index instead: 0 = one key per interval (.x offsets), 2 = one key per four
// intervals (.w offsets), used for chunks too small to fill the fine keys.
__device__ __forceinline__ int bfe_point_key(const EofGeom& ge, const SlGeom& gs, const int4* __restrict__ kc, int ncell, int subbits,
                                             double px, double py, double pz) {
    int cell;
    bfe_eof_cell_fast(ge, px, py, pz, cell);
    if ((unsigned)cell >= (unsigned)ncell) cell = 0;
    const float xf = (float)px, yf = (float)py, zf = (float)pz;
    const int j = bfe_sl_interval_fast(gs, sqrtf(fmaf(xf, xf, fmaf(yf, yf, zf * zf))));
    if (kc) {
        const int4 t = __ldg(kc + cell);
        int d = j - t.y;
        if (t.z <= 256) { d = d < 0 ? 0 : (d >= t.z ? t.z - 1 : d); }
        else d &= 255;
        return subbits ? (t.w + (d >> 2)) : (t.x + d);
    }
    return (cell << subbits) | (j & ((1 << subbits) - 1));
}

// radial-interval range of every table cell: the cell covers X in [ix, ix+1], Y in [iy, iy+1] of the table's coordinates; the
// radius over it runs from (R_lo, min |z|) to (R_hi, max |z|); cells on the table's rim are open (clamped bin indices)
__global__ void key_cell_span_kernel(EofGeom ge, SlGeom gs, const double* __restrict__ xi, int4* __restrict__ kc) {
    const int ncell = ge.numx * ge.numy;
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= ncell) return;
    const int ix = cell / ge.numy, iy = cell - ix * ge.numy;
    auto R_of = [&](double X) {                           // inverse of bfe_r_to_xi
        const double v = ge.xmin + X * ge.dx;
        if (ge.cmap == 1) return (v >= 1.0) ? 1.0e300 : ge.ascale * (1.0 + v) / (1.0 - v);
        if (ge.cmap == 2) return exp(v);
        return v;
    };
    auto z_of = [&](double Y) { return ge.hscale * sinh(ge.ymin + Y * ge.dy); };
    double Rlo = (ix == 0) ? 0.0 : fmax(R_of((double)ix), 0.0);
    double Rhi = fmax(R_of((double)ix + 1.0), 0.0);
    const double za = z_of((double)iy), zb = z_of((double)iy + 1.0);
    const double zmin = (za <= 0.0 && zb >= 0.0) ? 0.0 : fmin(fabs(za), fabs(zb));
    const double zmax = fmax(fabs(za), fabs(zb));
    const bool open = (ix == ge.numx - 1) || (iy == 0) || (iy == ge.numy - 1);
    const double rmin = sqrt(Rlo * Rlo + zmin * zmin), rmax = sqrt(Rhi * Rhi + zmax * zmax);
    int jmin = bfe_sl_bin(gs, xi, fmax(rmin, 1.0e-300)).i - 1;      // one interval of slack either side (FP32 keys)
    int jmax = (open || !(rmax < 1.0e299)) ? gs.numr - 2 : bfe_sl_bin(gs, xi, rmax).i + 1;
    if (jmin < 0) jmin = 0;
    if (jmax > gs.numr - 2) jmax = gs.numr - 2;
    if (jmax < jmin) jmax = jmin;
    const int span = jmax - jmin + 1;
    kc[cell] = make_int4(span <= 256 ? span : 256, jmin, span, 0);   // .x: keys of the cell, turned into the offset by the scan
}

// exclusive scans of the key counts over the cells (one CTA): fine keys (one per interval) -> .x, coarse keys (one per four
// intervals) -> .w; totals -> nkeys_out[0], nkeys_out[1]
__global__ void __launch_bounds__(1024) key_cell_scan_kernel(int ncell, int4* __restrict__ kc, int* __restrict__ nkeys_out) {
    __shared__ int s_w[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (ncell + 1023) / 1024, c0 = tid * per, c1 = min(ncell, c0 + per);
    int sum = 0, sum2 = 0;
    for (int c = c0; c < c1; ++c) { const int k = kc[c].x; sum += k; sum2 += (k + 3) >> 2; }
    int incl = sum, incl2 = sum2;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, off), v2 = __shfl_up_sync(0xffffffffu, incl2, off);
        if (lane >= off) { incl += v; incl2 += v2; }
    }
    if (lane == 31) { s_w[0][warp] = incl; s_w[1][warp] = incl2; }
    __syncthreads();
    if (warp == 0) {
        int w = s_w[0][lane], w2 = s_w[1][lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, w, off), v2 = __shfl_up_sync(0xffffffffu, w2, off);
            if (lane >= off) { w += v; w2 += v2; }
        }
        s_w[0][lane] = w; s_w[1][lane] = w2;
    }
    __syncthreads();
    int run = incl - sum + (warp > 0 ? s_w[0][warp - 1] : 0), run2 = incl2 - sum2 + (warp > 0 ? s_w[1][warp - 1] : 0);
    for (int c = c0; c < c1; ++c) {
        int4 t = kc[c];
        const int k = t.x;
        t.x = run; t.w = run2;
        kc[c] = t;
        run += k; run2 += (k + 3) >> 2;
    }
    if (tid == 1023) { nkeys_out[0] = s_w[0][31]; nkeys_out[1] = s_w[1][31]; }
}

// Guided tickets over the 128-point tiles of the sorted order: the first 3/4 of the tiles are handed out four at a time
// (a CTA that walks consecutive tiles finds the table lines of the previous tile in L1: with single-tile tickets every
// tile started cold, L1 hit rate 39-55 %), the rest one at a time so that the finish line is one tile wide.
// Small sets (fewer than 8 tiles per CTA of the grid: 1.25 x 10^5 orbits per GPU when 10^6 are spread over 8 GPUs) get single
// tiles only -- with four-tile tickets most CTAs took exactly one ticket and the kernel lasted as long as four tiles x K steps
// while the average CTA had work for 2.2 (C4 at N = 8: 56 us per step against 22 us at the single-GPU rate).
struct TileTicket { unsigned int first, count; };
__device__ __forceinline__ unsigned int bfe_ticket_n4(unsigned int ntile) {
    return (ntile >= 8u * gridDim.x) ? (ntile - ntile / 4) / 4 : 0u;
}
__device__ __forceinline__ unsigned int bfe_ticket_count(unsigned int ntile) {
    const unsigned int n4 = bfe_ticket_n4(ntile);
    return n4 + (ntile - 4 * n4);
}
__device__ __forceinline__ TileTicket bfe_ticket_tiles(unsigned int k, unsigned int ntile) {
    const unsigned int n4 = bfe_ticket_n4(ntile);
    TileTicket t;
    if (k < n4) { t.first = 4 * k; t.count = 4; }
    else { t.first = 4 * n4 + (k - n4); t.count = 1; }
    return t;
}

// rank of this lane's item inside its bin: lanes of the warp with the same key share ONE atomic
__device__ __forceinline__ int bfe_claim_rank(int* __restrict__ hist, int key, bool on) {
    const unsigned act = __ballot_sync(0xffffffffu, on);
    int rank = 0;
    if (on) {
        const unsigned peers = __match_any_sync(act, key);
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(hist + key, __popc(peers));
        base = __shfl_sync(peers, base, leader);
        rank = base + __popc(peers & ((1u << lane) - 1u));
    }
    return rank;
}

__global__ void __launch_bounds__(256)
pt_key_kernel(EofGeom ge, SlGeom gs, const int4* __restrict__ kc, int ncell, int subbits, int64_t n, const double* __restrict__ x,
              const double* __restrict__ y, const double* __restrict__ z, int* __restrict__ hist, int2* __restrict__ keyrank) {
    bfe_pdl_wait();
    bfe_pdl_trigger();
    // points arrive in the caller's order: the 32 keys of a warp differ, so one atomic per lane (no match / aggregation).  Four
    // points per thread per pass: the four rank claims (atomics WITH return, ~1 us each) are in flight together
    constexpr int U = 4;
    for (int64_t base = (int64_t)blockIdx.x * (blockDim.x * U) + threadIdx.x; base < n; base += (int64_t)gridDim.x * (blockDim.x * U)) {
        double px[U], py[U], pz[U];
        int key[U], rank[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            const bool on = i < n;
            px[u] = on ? __ldg(x + i) : 1.0; py[u] = on ? __ldg(y + i) : 0.0; pz[u] = on ? __ldg(z + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) key[u] = bfe_point_key(ge, gs, kc, ncell, subbits, px[u], py[u], pz[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) rank[u] = (base + (int64_t)u * blockDim.x < n) ? atomicAdd(hist + key[u], 1) : 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            if (i < n) keyrank[i] = make_int2(key[u], rank[u]);
        }
    }
}

// (Tried and removed: the rank claims aggregated per CTA in a shared-memory hash table, one global atomic per distinct key of a
// 2048-point pass, against the ~1 % of a chunk's points that share its hottest key.  pt_key 205 -> 176 us per 4 x 10^6 points
// standalone, C3 20.8 -> 21.3 ms in the pipeline: the kernel is bound by the RATE of random L2 requests -- per point one table
// look-up, one atomic -- not by same-address serialisation; ncu stall samples: 20 % on the coordinate loads, 34 % on the
// per-cell table look-up, 31 % on the atomic's return.  Staging the per-cell table in shared memory, 128 kB, one 1024-thread CTA per SM,
// did not change the kernel's time either, 207 vs 205 us: what is left is the rate of the atomics with return, ~21 G/s.)
// start[k] = exclusive prefix of hist inside its 1024-bin block, hist cleared; the last CTA turns the block totals into
// block prefixes.  Position of an item = bprefix[key >> 10] + start[key] + rank.
// 256 threads x 4 keys per block of 1024 keys: a block needs 256 x ~24 registers, so it fits on an SM beside the three resident
// CTAs of the evaluation kernel (the point path runs the support kernels on a second stream WHILE the evaluation runs).
__global__ void __launch_bounds__(256)
key_scan_kernel(int nkeys, int* __restrict__ hist, int* __restrict__ start, int* __restrict__ btot,
                int* __restrict__ bprefix, unsigned int* __restrict__ counter, int ticket_slot) {
    __shared__ int s_w[8];
    __shared__ bool s_last;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    if (blockIdx.x == 0 && threadIdx.x == 0) counter[1 + ticket_slot] = 0u;      // tile ticket of the evaluation kernel that follows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = blockIdx.x * BFE_SCAN_BLOCK + 4 * tid;                          // nkeys_cap is a multiple of 4: hist / start are 16-byte aligned
    int4 v = make_int4(0, 0, 0, 0);
    if (k0 + 3 < nkeys) {
        v = __ldcg(reinterpret_cast<const int4*>(hist + k0));
        *reinterpret_cast<int4*>(hist + k0) = make_int4(0, 0, 0, 0);
    } else {
        if (k0 < nkeys) { v.x = __ldcg(hist + k0); hist[k0] = 0; }
        if (k0 + 1 < nkeys) { v.y = __ldcg(hist + k0 + 1); hist[k0 + 1] = 0; }
        if (k0 + 2 < nkeys) { v.z = __ldcg(hist + k0 + 2); hist[k0 + 2] = 0; }
    }
    const int sum = v.x + v.y + v.z + v.w;
    int incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += u; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int wpre = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int t = s_w[q]; if (q < warp) wpre += t; tot += t; }
    const int excl = incl - sum + wpre;
    const int4 o = make_int4(excl, excl + v.x, excl + v.x + v.y, excl + v.x + v.y + v.z);
    if (k0 + 3 < nkeys) *reinterpret_cast<int4*>(start + k0) = o;
    else {
        if (k0 < nkeys) start[k0] = o.x;
        if (k0 + 1 < nkeys) start[k0 + 1] = o.y;
        if (k0 + 2 < nkeys) start[k0 + 2] = o.z;
    }
    if (tid == 0) btot[blockIdx.x] = tot;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // up to 2048 blocks (nkeys <= 2^21): eight totals per thread
    const int nb = gridDim.x;
    int t[8], s8 = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int bq = 8 * tid + q; t[q] = bq < nb ? __ldcg(btot + bq) : 0; s8 += t[q]; }
    int inc2 = s8;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc2, off); if (lane >= off) inc2 += u; }
    __syncthreads();
    if (lane == 31) s_w[warp] = inc2;
    __syncthreads();
    int wp2 = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (q < warp) wp2 += s_w[q];
    int run = inc2 - s8 + wp2;
#pragma unroll
    for (int q = 0; q < 8; ++q) { const int bq = 8 * tid + q; if (bq < nb) bprefix[bq] = run; run += t[q]; }
    if (tid == 0) *counter = 0u;
}

__global__ void __launch_bounds__(256)
perm_scatter_kernel(int64_t n, const int2* __restrict__ keyrank, const int* __restrict__ start,
                    const int* __restrict__ bprefix, int* __restrict__ perm) {
    bfe_pdl_wait();
    bfe_pdl_trigger();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int2 kr = __ldcs(keyrank + i);
        const int pos = __ldg(bprefix + (kr.x >> 10)) + __ldg(start + kr.x) + kr.y;
        perm[pos] = (int)i;
    }
}

// ---------------------------------------------------------------------------
// Fields.return_forces_cart / _cyl (potential.py:445-497 / 389-440) at the points of one chunk, in key order
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rec_scatter_kernel(int64_t n, const int2* __restrict__ keyrank, const int* __restrict__ start, const int* __restrict__ bprefix,
                   const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                   double* __restrict__ rec4, int* __restrict__ inv) {
    bfe_pdl_wait();
    bfe_pdl_trigger();
    // two points per thread per pass: their dependent position look-ups (bprefix, start) and coordinate loads overlap
    constexpr int U = 2;
    for (int64_t base = (int64_t)blockIdx.x * (blockDim.x * U) + threadIdx.x; base < n; base += (int64_t)gridDim.x * (blockDim.x * U)) {
        int2 kr[U];
        double px[U], py[U], pz[U];
        int pos[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            const bool on = i < n;
            kr[u] = on ? __ldcs(keyrank + i) : make_int2(0, 0);
            px[u] = on ? __ldg(x + i) : 0.0; py[u] = on ? __ldg(y + i) : 0.0; pz[u] = on ? __ldg(z + i) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) pos[u] = __ldg(bprefix + (kr[u].x >> 10)) + __ldg(start + kr[u].x) + kr[u].y;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            if (i < n) {
                bfe_st256(rec4 + 4 * (size_t)pos[u], px[u], py[u], pz[u], __longlong_as_double((long long)i));
                inv[i] = pos[u];
            }
        }
    }
}

// The point-path evaluation kernel is capped at 144 registers (no spills; under __launch_bounds__(128, 3) ptxas took 168 of the 170 it
// was allowed): 3 CTAs x 128 threads x 144 registers leave 10 240 registers of an SM free, which is one CTA of each support kernel
// (key 128 x 48, scan 256 x 24, scatter 256 x 32, gather 128 x 60) -- so the support stream runs WHILE the evaluation runs instead of
// between evaluation kernels.  BFE_EVAL_MAXNREG=0 restores the launch bounds.
#ifndef BFE_EVAL_MAXNREG
#define BFE_EVAL_MAXNREG 144
#endif
#if BFE_EVAL_MAXNREG > 0
#define BFE_FIELD_REC_QUAL __maxnreg__(BFE_EVAL_MAXNREG)
#else
#define BFE_FIELD_REC_QUAL __launch_bounds__(128, BFE_PERM_MINB)
#endif
template <int MCAP, int LCAP, bool CYL, bool F32>
__global__ void BFE_FIELD_REC_QUAL
field_rec_kernel(EofGeom ge, const void* __restrict__ G4, SlGeom gs, const void* __restrict__ A3,
                 const double* __restrict__ xi, const double* __restrict__ p0tab, const SlFacP fac,
                 int64_t n, const double* __restrict__ rec4, double crot, double srot, double* __restrict__ slot8,
                 unsigned int* __restrict__ ticket, int static_grid) {
    __shared__ unsigned int s_tile;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    const unsigned int ntile = (unsigned int)((n + 127) >> 7);
    const unsigned int nticket = bfe_ticket_count(ntile);
    for (;;) {
        TileTicket tk;
        if (static_grid) {
            // one CTA per four consecutive tiles, no persistent loop: SM slots come free every few tens of microseconds, so the
            // (high-priority) support kernels of the other stream get onto the SMs WHILE this kernel runs -- a persistent grid
            // at 3 x 128 x 168 registers fills the register file until the chunk is done and serialises the two streams
            const unsigned int per = (unsigned int)static_grid;            // tiles per CTA: 4, or 1 for small sets
            tk.first = per * blockIdx.x;
            tk.count = (tk.first + per <= ntile) ? per : (ntile > tk.first ? ntile - tk.first : 0u);
            if (tk.count == 0u) break;
        } else {
            __syncthreads();
            if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
            __syncthreads();
            if (s_tile >= nticket) break;
            tk = bfe_ticket_tiles(s_tile, ntile);
        }
        int64_t pos = (int64_t)tk.first * 128 + threadIdx.x;
        double px = 0.0, py = 0.0, pz = 0.0, id_ = 0.0;
        if (pos < n) bfe_ld256_nc(rec4 + 4 * (size_t)pos, px, py, pz, id_);
        for (unsigned int u = 0; u < tk.count; ++u) {
            const int64_t npos = pos + 128;
            double nx = 0.0, ny = 0.0, nz = 0.0;
            if (u + 1 < tk.count && npos < n) bfe_ld256_nc(rec4 + 4 * (size_t)npos, nx, ny, nz, id_);   // next tile's record in flight

            if (pos < n) {
                const CartForce f = bfe_field_cart_blk<MCAP, LCAP, CYL, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
                char* d = reinterpret_cast<char*>(slot8 + 8 * (size_t)pos);
                bfe_st256(d, f.fxd, f.fyd, f.fzd, f.pd);           // slot = {disc: fx, fy, fz, p | halo: fx, fy, fz, p}
                bfe_st256(d + 32, f.fxh, f.fyh, f.fzh, f.ph);
            }
            pos = npos; px = nx; py = ny; pz = nz;
        }
        if (static_grid) break;
    }
}

// ---------------------------------------------------------------------------
// Tried and removed (round 2): the same evaluation as TWO kernels, the disc half of a slot from the EOF blocks (80 registers,
// 6 CTAs per SM) and the halo half from the SL blocks (128 registers, 4 CTAs per SM), to get more warps per scheduler than the
// fused kernel's three.  No gain: 10^6 disc points 214 us against 212 fused; ncu (profiles/r02_ncu_full_field_half.csv,
// _fused.csv; 2^20 points): EOF half 70 us (long-scoreboard 11.6 warps per issue, l1tex 62 %, FP64 pipe 37 %) + SL half
// 97 us (FP64 pipe 64.5 %, math-pipe throttle 1.1) = 167 us against 161 us fused (FP64 pipe 50.5 %): the EOF phase is bound
// by its 42 x 8 L1 data-pipe wavefronts per point and the SL phase by FP64 issue, the fused kernel's time is already the sum
// of the two, and occupancy buys neither anything.  (Results were also not bit-identical to the caller-order kernel.)
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Shared-memory staging of the table blocks of a tile (FP64 tables): BASELINE north_star (b), "tables staged in shared
// memory (TMA bulk copies where the table fits)".  The tables do not fit, the blocks of a TILE do: the 128 points of a tile
// of the key order sit in a few neighbouring cells and radial intervals, whose blocks are CONTIGUOUS in G4 / A3.  Per tile
//   1. every thread finds its bins (bfe_field_prologue);
//   2. block-wide min / max of the cell and interval indices (warp redux + 4 x 4 values through shared memory);
//   3. one thread issues TWO TMA bulk copies (cp.async.bulk + mbarrier): blocks [cmin, cmin + NE) of G4 and
//      [jmin, jmin + NS) of A3, clipped to the tile's range -- at most (8 + 16) x 1344 B;
//   4. all threads wait on the mbarrier and evaluate (bfe_field_epilogue) from the staged copy -- warp-uniform 16-byte
//      shared-memory loads cost ~1.5 data-pipe cycles against ~8.3 for the 32-byte global load that hits L1, 2.8x less per
//      byte (profiles/probes/lds_broadcast_probe.cu), and the L1 data pipe is what bounds the per-lane kernels --
//      while a thread whose cell or interval lies outside the staged range (sparse outskirts, halo points with ~100 intervals
//      per cell) reads its block from global memory through the same generic loads.  Same arithmetic, same bits.
// The other resident CTAs of the SM cover the ~1 us a tile waits for its copies.
// ---------------------------------------------------------------------------
#define BFE_STAGE_NE 8
#define BFE_STAGE_NS 16
#define BFE_STAGE_BLK 84                       // double2 per block, both tables at mmax = 6 / lmax = 6 (1344 B)
#define BFE_STAGE_SMEM ((BFE_STAGE_NE + BFE_STAGE_NS) * BFE_STAGE_BLK * 16 + 256)

struct StageCtx {
    double2* sE; double2* sS;
    int* red;                                  // [2 parities][4 warps][4]
    unsigned int bar, parity, flip;            // mbarrier phase parity; flip: which half of `red` this tile uses
};

__device__ __forceinline__ StageCtx bfe_stage_init(unsigned char* smem) {
    StageCtx c;
    c.sE = reinterpret_cast<double2*>(smem);
    c.sS = c.sE + BFE_STAGE_NE * BFE_STAGE_BLK;
    unsigned char* tail = smem + (BFE_STAGE_NE + BFE_STAGE_NS) * BFE_STAGE_BLK * 16;
    c.red = reinterpret_cast<int*>(tail + 16);
    c.bar = bfe_smem_u32(tail);
    c.parity = 0u; c.flip = 0u;
    if (threadIdx.x == 0) {
        bfe_mbar_init(c.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    return c;
}

// all 128 threads of the CTA call this together; `on` = this thread has a point
template <int MCAP, int LCAP, bool CYL>
__device__ __forceinline__ CartForce bfe_stage_eval(const EofGeom& ge, const double2* __restrict__ G4, const SlGeom& gs,
                                                    const double2* __restrict__ A3, const double* __restrict__ p0tab,
                                                    const SlFacP& fac, const FieldPt& p, bool on, StageCtx& c) {
    constexpr int NPAIR = (LCAP + 1) * (LCAP + 2) / 2;
    constexpr int SBLK = BFE_A3_STRIDE(NPAIR);
    const int eblk = 12 * (ge.mmax + 1);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cell = p.eb.cell, j = (p.sb.i == 0) ? 1 : p.sb.i;
    const int c0 = __reduce_min_sync(0xffffffffu, on ? cell : 0x7fffffff), c1 = __reduce_max_sync(0xffffffffu, on ? cell : -1);
    const int j0 = __reduce_min_sync(0xffffffffu, on ? j : 0x7fffffff), j1 = __reduce_max_sync(0xffffffffu, on ? j : -1);
    int* red = c.red + 16 * (int)c.flip;       // double-buffered per tile: a fast warp never overwrites what a slow one still reads
    c.flip ^= 1u;
    if (lane == 0) { red[4 * warp] = c0; red[4 * warp + 1] = c1; red[4 * warp + 2] = j0; red[4 * warp + 3] = j1; }
    __syncthreads();                           // also: every thread has finished with the previous tile's staged blocks
    int cmin = red[0], cmax = red[1], jmin = red[2], jmax = red[3];
#pragma unroll
    for (int w = 1; w < 4; ++w) {
        cmin = min(cmin, red[4 * w]); cmax = max(cmax, red[4 * w + 1]);
        jmin = min(jmin, red[4 * w + 2]); jmax = max(jmax, red[4 * w + 3]);
    }
    const int ne = min(cmax - cmin + 1, BFE_STAGE_NE), ns = min(jmax - jmin + 1, BFE_STAGE_NS);
    CartForce f = {};
    if (cmax < cmin) return f;                 // no point in this tile (uniform across the CTA)
    if (threadIdx.x == 0) {
        const unsigned int be = (unsigned int)(ne * eblk * 16), bs = (unsigned int)(ns * SBLK * 16);
        bfe_mbar_expect_tx(c.bar, be + bs);
        bfe_bulk_g2s(bfe_smem_u32(c.sE), G4 + (size_t)cmin * eblk, be, c.bar);
        bfe_bulk_g2s(bfe_smem_u32(c.sS), A3 + (size_t)jmin * SBLK, bs, c.bar);
    }
    bfe_mbar_wait(c.bar, c.parity);
    c.parity ^= 1u;
    if (on) {
        const double2* baseE = (cell - cmin < ne) ? c.sE + (cell - cmin) * eblk : G4 + (size_t)cell * eblk;
        const double2* baseS = (j - jmin < ns) ? c.sS + (j - jmin) * SBLK : A3 + (size_t)j * SBLK;
        f = bfe_field_epilogue<MCAP, LCAP, CYL, SlFacP, LdGeneric>(ge, gs, baseE, baseS, p0tab, fac, p);
    }
    return f;
}

template <int MCAP, int LCAP, bool CYL>
__global__ void __launch_bounds__(128, BFE_PERM_MINB)
field_stage_kernel(EofGeom ge, const double2* __restrict__ G4, SlGeom gs, const double2* __restrict__ A3,
                   const double* __restrict__ xi, const double* __restrict__ p0tab, const SlFacP fac,
                   int64_t n, const double* __restrict__ rec4, double crot, double srot, double* __restrict__ slot8,
                   unsigned int* __restrict__ ticket) {
    extern __shared__ __align__(128) unsigned char s_stage[];
    __shared__ unsigned int s_tile;
    StageCtx ctx = bfe_stage_init(s_stage);
    bfe_pdl_wait();
    bfe_pdl_trigger();
    const unsigned int ntile = (unsigned int)((n + 127) >> 7);
    const unsigned int nticket = bfe_ticket_count(ntile);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        if (s_tile >= nticket) break;
        const TileTicket tk = bfe_ticket_tiles(s_tile, ntile);
        int64_t pos = (int64_t)tk.first * 128 + threadIdx.x;
        double px = 0.0, py = 0.0, pz = 0.0, id_ = 0.0;
        if (pos < n) bfe_ld256_nc(rec4 + 4 * (size_t)pos, px, py, pz, id_);
        for (unsigned int u = 0; u < tk.count; ++u) {
            const int64_t npos = pos + 128;
            double nx = 0.0, ny = 0.0, nz = 0.0;
            if (u + 1 < tk.count && npos < n) bfe_ld256_nc(rec4 + 4 * (size_t)npos, nx, ny, nz, id_);   // next tile's record in flight
            const bool on = pos < n;
            FieldPt p = {};
            if (on) p = bfe_field_prologue<CYL>(ge, gs, xi, px, py, pz, crot, srot);
            const CartForce f = bfe_stage_eval<MCAP, LCAP, CYL>(ge, G4, gs, A3, p0tab, fac, p, on, ctx);
            if (on) {
                char* d = reinterpret_cast<char*>(slot8 + 8 * (size_t)pos);
                bfe_st256(d, f.fxd, f.fyd, f.fzd, f.pd);           // slot = {disc: fx, fy, fz, p | halo: fx, fy, fz, p}
                bfe_st256(d + 32, f.fxh, f.fyh, f.fzh, f.ph);
            }
            pos = npos; px = nx; py = ny; pz = nz;
        }
    }
}

__global__ void __launch_bounds__(256)
field_gather_kernel(int64_t n, int64_t ntot, const int* __restrict__ inv, const double* __restrict__ slot8,
                    double* __restrict__ out8) {
    bfe_pdl_wait();
    bfe_pdl_trigger();
    // two points per thread per pass: four random 32-byte slot reads in flight
    constexpr int U = 2;
    for (int64_t base = (int64_t)blockIdx.x * (blockDim.x * U) + threadIdx.x; base < n; base += (int64_t)gridDim.x * (blockDim.x * U)) {
        double a[U][4], b[U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            const char* s = reinterpret_cast<const char*>(slot8 + 8 * (size_t)(i < n ? __ldg(inv + i) : 0));
            bfe_ld256_nc(s, a[u][0], a[u][1], a[u][2], a[u][3]);
            bfe_ld256_nc(s + 32, b[u][0], b[u][1], b[u][2], b[u][3]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + (int64_t)u * blockDim.x;
            if (i < n) {
                // slot = {disc: fx, fy, fz, p | halo: fx, fy, fz, p}; output rows fxd, fxh, fyd, fyh, fzd, fzh, pd, ph
                out8[i] = a[u][0]; out8[2 * ntot + i] = a[u][1]; out8[4 * ntot + i] = a[u][2]; out8[6 * ntot + i] = a[u][3];
                out8[ntot + i] = b[u][0]; out8[3 * ntot + i] = b[u][1]; out8[5 * ntot + i] = b[u][2]; out8[7 * ntot + i] = b[u][3];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// integrate.leapfrog_integrate (integrate.py:53-190) on orbit records that stay in the caller's order
// ---------------------------------------------------------------------------
struct __align__(32) OrbRec {            // 96 bytes = three full sectors
    double x, y, z, vx;
    double vy, vz, dt, ax;
    double ay, az, pad0, pad1;
};

__global__ void __launch_bounds__(256)
orbit_pack_key_kernel(EofGeom ge, SlGeom gs, const int4* __restrict__ kc, int ncell, int subbits, int64_t n, const double* __restrict__ state6,
                      double dt, const double* __restrict__ dt_orbit, OrbRec* __restrict__ rec, int* __restrict__ hist,
                      int2* __restrict__ keyrank) {
    bfe_pdl_wait();
    bfe_pdl_trigger();
    const int64_t nround = (n + 31) & ~(int64_t)31;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += (int64_t)gridDim.x * blockDim.x) {
        const bool on = i < n;
        int key = 0;
        if (on) {
            const double px = state6[i], py = state6[n + i], pz = state6[2 * n + i];
            char* dst = reinterpret_cast<char*>(rec + i);
            bfe_st256(dst, px, py, pz, state6[3 * n + i]);
            bfe_st256(dst + 32, state6[4 * n + i], state6[5 * n + i], dt_orbit ? dt_orbit[i] : dt, 0.0);
            bfe_st256(dst + 64, 0.0, 0.0, 0.0, 0.0);
            key = bfe_point_key(ge, gs, kc, ncell, subbits, px, py, pz);
        }
        const int rank = bfe_claim_rank(hist, key, on);
        if (on) keyrank[i] = make_int2(key, rank);
    }
}

// nsteps velocity-Verlet steps step0+1 .. step0+nsteps of every orbit, in key order.  first: the acceleration at step0 is
// evaluated here (otherwise it is the one the previous launch left in the record: the same bits).  rekey: key and rank of
// the end position for the next re-sort.
template <int MCAP, int LCAP, bool F32>
__global__ void __launch_bounds__(128, BFE_PERM_MINB)
leapfrog_perm_kernel(EofGeom ge, const void* __restrict__ G4, SlGeom gs, const void* __restrict__ A3,
                     const double* __restrict__ xi, const double* __restrict__ p0tab, const SlFacP fac,
                     int64_t norbit, int64_t step0, int nsteps, double rotfreq, int first, int rekey, const int4* __restrict__ kc, int ncell, int subbits,
                     const int* __restrict__ perm, OrbRec* __restrict__ rec, int* __restrict__ hist,
                     int2* __restrict__ keyrank, unsigned int* __restrict__ ticket) {
    __shared__ unsigned int s_tile;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    const unsigned int ntile = (unsigned int)((norbit + 127) >> 7);
    const unsigned int nticket = bfe_ticket_count(ntile);
    const double w = BFE_TWOPI * rotfreq;                    // barpos = 2 pi rotfreq (k dt), integrate.py:94-97
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        if (s_tile >= nticket) break;
        const TileTicket tk = bfe_ticket_tiles(s_tile, ntile);
        int64_t pos = (int64_t)tk.first * 128 + threadIdx.x;
        int i = pos < norbit ? __ldg(perm + pos) : -1;
        for (unsigned int u = 0; u < tk.count; ++u) {
            const int64_t npos = pos + 128;
            const int ni = (u + 1 < tk.count && npos < norbit) ? __ldg(perm + npos) : -1;
            int key = 0;
            if (i >= 0) {
                double px, py, pz, vx, vy, vz, dt, ax, ay, az, u0, u1;
                {
                    // plain (coherent) loads: the record is rewritten in place by this kernel
                    const double4* r4 = reinterpret_cast<const double4*>(rec + i);
                    const double4 a = r4[0], b = r4[1], c = r4[2];
                    px = a.x; py = a.y; pz = a.z; vx = a.w; vy = b.x; vz = b.y; dt = b.z; ax = b.w; ay = c.x; az = c.y;
                    u0 = c.z; u1 = c.w;
                }
                const double hdt2 = 0.5 * (dt * dt);
                double srot, crot;
                if (first) {
                    sincos(w * ((double)step0 * dt), &srot, &crot);
                    const CartForce f = bfe_field_cart_blk<MCAP, LCAP, false, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
                    ax = f.fxd + f.fxh; ay = f.fyd + f.fyh; az = f.fzd + f.fzh;
                }
                for (int k = 1; k <= nsteps; ++k) {
                    const int64_t step = step0 + k;
                    px = px + (vx * dt) + (ax * hdt2);                   // integrate.py:129-131
                    py = py + (vy * dt) + (ay * hdt2);
                    pz = pz + (vz * dt) + (az * hdt2);
                    sincos(w * ((double)step * dt), &srot, &crot);
                    const CartForce f = bfe_field_cart_blk<MCAP, LCAP, false, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
                    const double bx = f.fxd + f.fxh, by = f.fyd + f.fyh, bz = f.fzd + f.fzh;             // 134-138
                    vx = vx + (0.5 * (ax + bx) * dt);                    // 141-143
                    vy = vy + (0.5 * (ay + by) * dt);
                    vz = vz + (0.5 * (az + bz) * dt);
                    ax = bx; ay = by; az = bz;
                }
                char* d = reinterpret_cast<char*>(rec + i);
                bfe_st256(d, px, py, pz, vx);
                bfe_st256(d + 32, vy, vz, dt, ax);
                bfe_st256(d + 64, ay, az, u0, u1);
                // key of the NEXT evaluation point: the drift of the following step (the same expression, so the same bits)
                // -- orbits cross an SL interval per step, so a key taken at the current position would be stale at once
                if (rekey) key = bfe_point_key(ge, gs, kc, ncell, subbits, px + (vx * dt) + (ax * hdt2), py + (vy * dt) + (ay * hdt2),
                                               pz + (vz * dt) + (az * hdt2));
            }
            if (rekey) {
                const int rank = bfe_claim_rank(hist, key, i >= 0);
                if (i >= 0) keyrank[i] = make_int2(key, rank);
            }
            i = ni; pos = npos;
        }
    }
}

// leapfrog_perm_kernel with the table blocks of every step's tile staged in shared memory (bfe_stage_eval), FP64 tables
template <int MCAP, int LCAP>
__global__ void __launch_bounds__(128, BFE_PERM_MINB)
leapfrog_stage_kernel(EofGeom ge, const double2* __restrict__ G4, SlGeom gs, const double2* __restrict__ A3,
                      const double* __restrict__ xi, const double* __restrict__ p0tab, const SlFacP fac,
                      int64_t norbit, int64_t step0, int nsteps, double rotfreq, int first, int rekey, const int4* __restrict__ kc, int ncell, int subbits,
                      const int* __restrict__ perm, OrbRec* __restrict__ rec, int* __restrict__ hist,
                      int2* __restrict__ keyrank, unsigned int* __restrict__ ticket) {
    extern __shared__ __align__(128) unsigned char s_stage[];
    __shared__ unsigned int s_tile;
    StageCtx ctx = bfe_stage_init(s_stage);
    bfe_pdl_wait();
    bfe_pdl_trigger();
    const unsigned int ntile = (unsigned int)((norbit + 127) >> 7);
    const unsigned int nticket = bfe_ticket_count(ntile);
    const double w = BFE_TWOPI * rotfreq;                    // barpos = 2 pi rotfreq (k dt), integrate.py:94-97
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        if (s_tile >= nticket) break;
        const TileTicket tk = bfe_ticket_tiles(s_tile, ntile);
        int64_t pos = (int64_t)tk.first * 128 + threadIdx.x;
        int i = pos < norbit ? __ldg(perm + pos) : -1;
        for (unsigned int u = 0; u < tk.count; ++u) {
            const int64_t npos = pos + 128;
            const int ni = (u + 1 < tk.count && npos < norbit) ? __ldg(perm + npos) : -1;
            const bool on = i >= 0;
            int key = 0;
            double px = 0.0, py = 0.0, pz = 0.0, vx = 0.0, vy = 0.0, vz = 0.0, dt = 0.0, ax = 0.0, ay = 0.0, az = 0.0, u0 = 0.0, u1 = 0.0;
            if (on) {
                // plain (coherent) loads: the record is rewritten in place by this kernel
                const double4* r4 = reinterpret_cast<const double4*>(rec + i);
                const double4 a = r4[0], b = r4[1], c = r4[2];
                px = a.x; py = a.y; pz = a.z; vx = a.w; vy = b.x; vz = b.y; dt = b.z; ax = b.w; ay = c.x; az = c.y;
                u0 = c.z; u1 = c.w;
            }
            const double hdt2 = 0.5 * (dt * dt);
            double srot, crot;
            if (first) {                                             // CTA-uniform
                FieldPt p = {};
                if (on) {
                    sincos(w * ((double)step0 * dt), &srot, &crot);
                    p = bfe_field_prologue<false>(ge, gs, xi, px, py, pz, crot, srot);
                }
                const CartForce f = bfe_stage_eval<MCAP, LCAP, false>(ge, G4, gs, A3, p0tab, fac, p, on, ctx);
                if (on) { ax = f.fxd + f.fxh; ay = f.fyd + f.fyh; az = f.fzd + f.fzh; }
            }
            for (int k = 1; k <= nsteps; ++k) {
                const int64_t step = step0 + k;
                FieldPt p = {};
                if (on) {
                    px = px + (vx * dt) + (ax * hdt2);                   // integrate.py:129-131
                    py = py + (vy * dt) + (ay * hdt2);
                    pz = pz + (vz * dt) + (az * hdt2);
                    sincos(w * ((double)step * dt), &srot, &crot);
                    p = bfe_field_prologue<false>(ge, gs, xi, px, py, pz, crot, srot);
                }
                const CartForce f = bfe_stage_eval<MCAP, LCAP, false>(ge, G4, gs, A3, p0tab, fac, p, on, ctx);
                if (on) {
                    const double bx = f.fxd + f.fxh, by = f.fyd + f.fyh, bz = f.fzd + f.fzh;         // 134-138
                    vx = vx + (0.5 * (ax + bx) * dt);                    // 141-143
                    vy = vy + (0.5 * (ay + by) * dt);
                    vz = vz + (0.5 * (az + bz) * dt);
                    ax = bx; ay = by; az = bz;
                }
            }
            if (on) {
                char* d = reinterpret_cast<char*>(rec + i);
                bfe_st256(d, px, py, pz, vx);
                bfe_st256(d + 32, vy, vz, dt, ax);
                bfe_st256(d + 64, ay, az, u0, u1);
                // key of the NEXT evaluation point: the drift of the following step (the same expression, so the same bits)
                if (rekey) key = bfe_point_key(ge, gs, kc, ncell, subbits, px + (vx * dt) + (ax * hdt2), py + (vy * dt) + (ay * hdt2),
                                               pz + (vz * dt) + (az * hdt2));
            }
            if (rekey) {
                const int rank = bfe_claim_rank(hist, key, on);
                if (on) keyrank[i] = make_int2(key, rank);
            }
            i = ni; pos = npos;
        }
    }
}

__global__ void __launch_bounds__(256)
orbit_unpack_kernel(int64_t n, const OrbRec* __restrict__ rec, double* __restrict__ state6, int* __restrict__ nsteps_out,
                    int nint) {
    bfe_pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a0, a1, a2, a3, b0, b1, b2, b3;
        bfe_ld256_nc(rec + i, a0, a1, a2, a3);
        bfe_ld256_nc(reinterpret_cast<const char*>(rec + i) + 32, b0, b1, b2, b3);
        state6[i] = a0; state6[n + i] = a1; state6[2 * n + i] = a2;
        state6[3 * n + i] = a3; state6[4 * n + i] = b0; state6[5 * n + i] = b1;
        if (nsteps_out) nsteps_out[i] = nint;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int g_bfe_orbit_resort = 3;             // option "orbit_resort": steps between re-sorts of a large orbit batch (0: plain kernel)
// options "key_subbits" (points) / "orbit_key_subbits" (orbits): low bits of the SL interval in the ordering key (bins = ncell << bits).
// Measured on 4e7-point sets (profiles/r02_chunk_probe2.json, us per 1e6 at 2^22-point chunks, bits 4 / 5 / 6 / 7 / 8): halo-like
// 342 / 322 / 310 / 271 / 265 (a cell of the outer halo spans > 100 radial intervals: 16 slots alias them), disc-like 189 / 186 / 189 /
// 198 / 207 (more bins = fewer points per bin and a longer scan).  Points: at most 7, fewer for small chunks (keysort_ws); orbits
// (disc-like batches, one sort per K steps): 4.
int g_bfe_key_mode = 1;                 // option "key_mode": bit 0 = points, bit 1 = orbits ordered by per-cell interval spans (bfe_point_key);
                                        // otherwise (cell << bits) | (interval & mask).  Orbits: measured 2.5 % slower with the spans (C4)
int g_bfe_field_gather_stream = 0;      // option "field_gather_stream": gather kernel of the point path on its own support stream
int g_bfe_field_support_slim = 1;       // option "field_support_slim": support kernels of the point path as one small CTA per SM beside the evaluation
int g_bfe_field_eval_static = 1;        // option "field_eval_static": one evaluation CTA per four tiles (1) instead of a persistent grid with tickets (0).
                                        // C3 21.75 -> 20.8 ms: the support stream's CTAs get SM slots while the evaluation runs.  (Evaluation limited
                                        // to 2 persistent CTAs per SM so that a third of the register file stays free: 27.1 ms -- it needs its warps.)
int g_bfe_keycell_nkeys_last = 0;       // read-only option "keycell_nkeys": keys of the last per-cell span table built
int g_bfe_key_subbits = 7;
int g_bfe_orbit_key_subbits = BFE_KEY_SUBBITS;
int g_bfe_orbit_sort_min = 65536;       // option "orbit_sort_min": smallest batch on the key-ordered path
// option "field_sort_chunk": most points per sort + evaluate pass.  2^22 since the probe above (2^20 / 2^21 / 2^22 / 2^23 / 2^24: halo
// 353 / 347 / 342 / 357 / 366, disc 205 / 196 / 189 / 206 / 212 at 4 bits): more points per bin beat L2 residency of the records;
// sets smaller than four such chunks are cut in four (>= 2^20 points each) so that two chunks are still in flight
int g_bfe_field_sort_chunk = 1 << 22;
int g_bfe_stage_eval = 0;               // option "stage_eval": FP64-table key-ordered kernels evaluate from TMA-staged shared-memory blocks (1) or
                                        // with global loads (0, default).  Measured on B200 (profiles/r02_field_probe_stage*.json, r02_ncu_full_field_stage.csv):
                                        // global-load requests per tile fall 3.7x and results stay bit-identical, but the warps of a 2^19..2^20-point
                                        // chunk hold ~3 distinct (cell, interval) pairs, so a 16-byte shared load still takes 3.7 data-pipe wavefronts
                                        // (1.5 when warp-uniform): 207 vs 210 us per 1e6 disc points, 387 vs 339 for halo points (their intervals
                                        // overflow the 16 staged blocks and fall back to twice as many 128-bit loads), orbits 0.191 vs 0.187 ns
int g_bfe_field_sort_min = 65536;       // option "field_sort_min": smallest point set evaluated in key order (0: never)

static size_t os_align(size_t v) { return (v + 255) / 256 * 256; }

struct KeySortWs {
    int ncell, subbits, nkeys, nblk;
    const int4* kc;                  // per-cell key spans (option key_mode = 1) or nullptr (bit scheme)
    int* hist; int* start; int* btot; int* bprefix; unsigned int* counter;     // counter[0]: scan's last-CTA count, [1], [2]: tile tickets
    int2* keyrank[2]; int* perm[2];  // perm doubles as the inverse permutation of the point path
    double* rec4[2]; double* slot8[2];   // point path only, two chunks in flight; slot8 aliases keyrank (dead after the scatter)
};

// workspace in he->orbit_ws, grown on demand: header + per item 12 bytes (orbits: keyrank, perm) or 2 x 100 bytes (points)
// per-cell key spans for the (EOF table, SL table) pair: built on first use and whenever another SL table is paired
static int key_cells(bfe_eof* he, bfe_sl* hs, cudaStream_t stream) {
    const int ncell = he->g.numx * he->g.numy;
    unsigned long long tag = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { tag = (tag ^ v) * 1099511628211ull; };
    unsigned long long u;
    mix((unsigned long long)hs->g.numr); mix((unsigned long long)hs->g.cmap);
    memcpy(&u, &hs->g.xi0, 8); mix(u); memcpy(&u, &hs->g.dxi, 8); mix(u); memcpy(&u, &hs->g.scale, 8); mix(u);
    mix((unsigned long long)(uintptr_t)hs->xi);
    if (tag == 0) tag = 1;
    if (he->keycell && he->keycell_tag == tag) return BFE_OK;
    if (!he->keycell) BFE_CUDA(cudaMalloc(&he->keycell, sizeof(int4) * (size_t)ncell + 16));
    int4* kc = (int4*)he->keycell;
    int* d_n = (int*)(kc + ncell);
    key_cell_span_kernel<<<(ncell + 255) / 256, 256, 0, stream>>>(he->g, hs->g, hs->xi, kc);
    BFE_LAUNCH_CHECK("key_cell_span_kernel");
    key_cell_scan_kernel<<<1, 1024, 0, stream>>>(ncell, kc, d_n);
    BFE_LAUNCH_CHECK("key_cell_scan_kernel");
    int nk[2] = {0, 0};
    BFE_CUDA(cudaMemcpyAsync(nk, d_n, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
    BFE_CUDA(cudaStreamSynchronize(stream));
    he->keycell_nkeys = nk[0];
    he->keycell_nkeys2 = nk[1];
    he->keycell_tag = tag;
    g_bfe_keycell_nkeys_last = nk[0];
    return BFE_OK;
}

static int keysort_ws(bfe_eof* he, bfe_sl* hs, cudaStream_t stream, int64_t cap_need, bool points, KeySortWs& w) {
    w.ncell = he->g.numx * he->g.numy;
    w.kc = nullptr;
    int nkeys_cells = 0, kc_shift = 0;
    if (points ? (g_bfe_key_mode & 1) : (g_bfe_key_mode & 2)) {
        int rc = key_cells(he, hs, stream);
        if (rc != BFE_OK) return rc;
        if (he->keycell_nkeys > 0 && he->keycell_nkeys <= (1 << 21)) {
            w.kc = (const int4*)he->keycell;
            // fine keys (one per interval) when the chunk brings >= ~5 points per key, else one key per four intervals
            // (2 x 10^6 disc points in 10^6-point chunks: 228 us per 10^6 with the fine keys, 214 with 5 bits of the bit scheme)
            kc_shift = (cap_need >= 5 * (int64_t)he->keycell_nkeys) ? 0 : 2;
            nkeys_cells = kc_shift ? he->keycell_nkeys2 : he->keycell_nkeys;
        }
    }
    auto clamp_bits = [&](int b) {
        if (b < 0) b = 0;
        if (b > 8) b = 8;
        while (b > 0 && ((int64_t)w.ncell << b) > ((int64_t)1 << 21)) --b;
        return b;
    };
    // points and orbits have their own key width (options key_subbits / orbit_key_subbits); the header is laid out for the
    // wider of the two, so alternating between the two paths does not re-lay the workspace
    const int sub_pts = clamp_bits(g_bfe_key_subbits), sub_orb = clamp_bits(g_bfe_orbit_key_subbits);
    w.subbits = points ? sub_pts : sub_orb;
    if (points) {
        // the key width follows the chunk: ~4 points per (cell, slot) on average -- 2^22-point chunks on the 128 x 64 table take
        // 7 bits, 2^19-point chunks 4 (with 7 bits a 2 x 10^6-point disc set lost its coherence: 303 us per 10^6 against 209)
        int lg = 0;
        while (((int64_t)w.ncell << (lg + 1)) <= cap_need) ++lg;             // floor(log2(chunk / ncell))
        int want = lg - 2;
        if (want < 4) want = 4;
        if (w.subbits > want) w.subbits = want;
    }
    const int sub_cap = sub_pts > sub_orb ? sub_pts : sub_orb;
    if (((int64_t)w.ncell << sub_cap) > ((int64_t)1 << 21)) return BFE_ERR_UNSUPPORTED;
    w.nkeys = w.kc ? nkeys_cells : (w.ncell << w.subbits);
    w.nblk = (w.nkeys + BFE_SCAN_BLOCK - 1) / BFE_SCAN_BLOCK;
    int nkeys_cap = w.ncell << sub_cap;
    if (w.kc && he->keycell_nkeys > nkeys_cap) nkeys_cap = he->keycell_nkeys;      // laid out for the fine keys: no re-lay when the shift changes
    if (w.kc) w.subbits = kc_shift;                    // with the per-cell table the kernels' `subbits` argument is the slot shift
    const int nblk_cap = (nkeys_cap + BFE_SCAN_BLOCK - 1) / BFE_SCAN_BLOCK;
    const size_t o_hist = 0;
    const size_t o_start = os_align(o_hist + sizeof(int) * (size_t)nkeys_cap);
    const size_t o_btot = os_align(o_start + sizeof(int) * (size_t)nkeys_cap);
    const size_t o_bpre = os_align(o_btot + sizeof(int) * (size_t)nblk_cap);
    const size_t o_ctr = os_align(o_bpre + sizeof(int) * (size_t)(nblk_cap + 1));
    const size_t o_item = os_align(o_ctr + 64);
    const size_t a64 = os_align(64 * (size_t)cap_need), a32 = os_align(32 * (size_t)cap_need), a8 = os_align(8 * (size_t)cap_need),
                 a4 = os_align(4 * (size_t)cap_need);
    const int64_t need = points ? (int64_t)(2 * (a64 + a32 + a4)) : (int64_t)(a8 + a4);
    // orbit_cap: BYTES available for items; orbit_hdr: bytes of the header the workspace was laid out for.  The header depends
    // on the option key_subbits (nkeys = ncell << subbits): a change of the option after the first call must re-lay the
    // workspace (it used to keep the old allocation and the 16x larger histogram of subbits 8 ran into the item arrays)
    if (need > he->orbit_cap || (int64_t)o_item != he->orbit_hdr || !he->orbit_ws) {
        const int64_t keep = he->orbit_cap > need ? he->orbit_cap : need + need / 8 + 4096;
        if (he->orbit_ws) { BFE_CUDA(cudaDeviceSynchronize()); BFE_CUDA(cudaFree(he->orbit_ws)); he->orbit_ws = nullptr; he->orbit_cap = 0; he->orbit_hdr = 0; }
        const int64_t bytes = keep;
        BFE_CUDA(cudaMalloc(&he->orbit_ws, o_item + (size_t)bytes));
        BFE_CUDA(cudaMemset(he->orbit_ws, 0, o_item));
        BFE_CUDA(cudaDeviceSynchronize());
        he->orbit_cap = bytes;
        he->orbit_hdr = (int64_t)o_item;
    }
    char* b = (char*)he->orbit_ws;
    w.hist = (int*)(b + o_hist); w.start = (int*)(b + o_start); w.btot = (int*)(b + o_btot); w.bprefix = (int*)(b + o_bpre);
    w.counter = (unsigned int*)(b + o_ctr);
    char* it = b + o_item;
    if (points) {
        for (int k = 0; k < 2; ++k) {
            char* q = it + (size_t)k * (a64 + a32 + a4);
            w.slot8[k] = (double*)q;  w.keyrank[k] = (int2*)q;
            w.rec4[k] = (double*)(q + a64);
            w.perm[k] = (int*)(q + a64 + a32);
        }
    } else {
        w.keyrank[0] = w.keyrank[1] = (int2*)it;
        w.perm[0] = w.perm[1] = (int*)(it + a8);
        w.rec4[0] = w.rec4[1] = nullptr; w.slot8[0] = w.slot8[1] = nullptr;
    }
    return BFE_OK;
}

// aux stream + events of the two-stream point pipeline (lazily made, owned by the EOF handle)
struct FieldPipe { cudaStream_t aux, aux2; cudaEvent_t ev_in, ev_k[2], ev_e[2], ev_g[2], ev_out, ev_out2; };

void bfe_field_pipe_destroy(void* p_) {
    FieldPipe* p = (FieldPipe*)p_;
    if (!p) return;
    cudaEventDestroy(p->ev_in); cudaEventDestroy(p->ev_out); cudaEventDestroy(p->ev_out2);
    for (int k = 0; k < 2; ++k) { cudaEventDestroy(p->ev_k[k]); cudaEventDestroy(p->ev_e[k]); cudaEventDestroy(p->ev_g[k]); }
    cudaStreamDestroy(p->aux); cudaStreamDestroy(p->aux2);
    delete p;
}

static int field_pipe(bfe_eof* he, FieldPipe*& out) {
    if (!he->field_pipe) {
        FieldPipe* p = new FieldPipe();
        // the support stream has the highest priority: its short CTAs take the SM slots the evaluation kernel frees first
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        cudaError_t e = cudaStreamCreateWithPriority(&p->aux, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&p->aux2, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_out2, cudaEventDisableTiming);
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
            e = cudaEventCreateWithFlags(&p->ev_k[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_e[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_g[k], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) { bfe_set_cuda_error(e, "field_pipe"); delete p; return BFE_ERR_CUDA; }
        he->field_pipe = p;
    }
    out = (FieldPipe*)he->field_pipe;
    return BFE_OK;
}

static int grid_cap(int64_t n, int block, int cap) {
    int64_t g = (n + block - 1) / block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

#define KS_LAUNCH(name, kern, grid, block, ...) KS_LAUNCH_ON(stream, name, kern, grid, block, __VA_ARGS__)
#define KS_LAUNCH_ON(strm, name, kern, grid, block, ...)                                                  \
    do {                                                                                                  \
        cudaError_t _e = bfe_launch(kern, dim3(grid), dim3(block), 0, strm, nullptr, 0, __VA_ARGS__);     \
        if (_e != cudaSuccess) { bfe_set_cuda_error(_e, name); return BFE_ERR_CUDA; }                     \
        bfe_count_launch(1);                                                                              \
    } while (0)

// Two chunks in flight: the caller's stream runs the FP64-bound evaluation kernels (three CTAs per SM, so a quarter of
// the register file stays free), an internal stream runs the latency-bound support kernels of the neighbouring chunks
// (key + scan + scatter of chunk c+1, gather of chunk c-1) underneath them; events order the two (serialised, the
// support kernels were 30 % of the pass: ncu launch list profiles/r02_launches_points_*.csv).
int bfe_field_force_sorted(bfe_eof* he, bfe_sl* hs, int64_t n, const double* x, const double* y, const double* z,
                           double crot, double srot, double* out8, bool cyl, cudaStream_t stream) {
    const int64_t chunk_opt = g_bfe_field_sort_chunk > 0 ? g_bfe_field_sort_chunk : (1 << 19);
    int64_t chunk = ((n + 3) / 4 + 127) / 128 * 128;
    if (chunk < ((int64_t)1 << 20)) chunk = (int64_t)1 << 20;
    if (chunk > chunk_opt) chunk = chunk_opt;
    if (chunk > n) chunk = n;
    KeySortWs w;
    int rc = keysort_ws(he, hs, stream, chunk, true, w);
    if (rc != BFE_OK) return rc;
    FieldPipe* fp = nullptr;
    rc = field_pipe(he, fp);
    if (rc != BFE_OK) return rc;
    const bool f32 = bfe_use_fp32(he);
    rc = f32 ? bfe_eof_ensure_g4f(he, stream) : bfe_eof_ensure_g4(he, stream);
    // FP64: the per-lane kernels read the per-interval polynomial blocks A4, the TMA-staged ones the three-node blocks A3
    const bool staged = !f32 && g_bfe_stage_eval && he->g.mmax <= 6;
    if (rc == BFE_OK) rc = f32 ? bfe_sl_ensure_a3f(hs, stream) : (staged ? bfe_sl_ensure_a3(hs, stream) : bfe_sl_ensure_a4(hs, stream));
    if (rc != BFE_OK) return rc;
    const void* G4 = f32 ? (const void*)he->g4f : (const void*)he->g4;
    const void* A3 = f32 ? (const void*)hs->a3f : (staged ? (const void*)hs->a3 : (const void*)hs->a4);
    const SlFacP facp = bfe_sl_facp(hs);
    const int kt = bfe_kt_begin("field_sorted_pass", stream);
    cudaStream_t aux = fp->aux;
    // option field_gather_stream: the gather of chunk c-1 on its own stream, beside key + scan + scatter of chunk c+1
    const bool gsep = g_bfe_field_gather_stream != 0;
    cudaStream_t gst = gsep ? fp->aux2 : fp->aux;
    BFE_CUDA(cudaEventRecord(fp->ev_in, stream));
    BFE_CUDA(cudaStreamWaitEvent(aux, fp->ev_in, 0));
    const int64_t nchunk = (n + chunk - 1) / chunk;
    auto sort_chunk = [&](int64_t c) -> int {
        const int64_t c0 = c * chunk, m = (n - c0) < chunk ? (n - c0) : chunk;
        const int b = (int)(c & 1);
        // the buffers of chunk c (keys = result slots, permutation) are those of chunk c-2: its gather (other support stream) must be done
        if (gsep && c >= 2) BFE_CUDA(cudaStreamWaitEvent(aux, fp->ev_g[b], 0));
        // slim support grids (option field_support_slim): one small CTA per SM, sized to fit beside the evaluation kernel's three
        // slim only when the set has >= 3 chunks: with one or two there is nothing to hide the slim kernels behind
        // (2 x 10^6 disc points: 246 us per 10^6 slim, 215 with full-size grids)
        const bool slim = g_bfe_field_support_slim != 0 && nchunk >= 3;
        const int sc = g_bfe_field_support_slim > 1 ? g_bfe_field_support_slim : 1;      // support CTAs per SM in slim mode
        const int g256 = slim ? grid_cap(m, 256, he->num_sms * sc) : grid_cap(m, 256, he->num_sms * 4);
        const int gkey = slim ? grid_cap(m, 128, he->num_sms * sc) : g256, bkey = slim ? 128 : 256;
        KS_LAUNCH_ON(aux, "pt_key_kernel", pt_key_kernel, gkey, bkey, he->g, hs->g, w.kc, w.ncell, w.subbits, m, x + c0, y + c0, z + c0,
                     w.hist, w.keyrank[b]);
        KS_LAUNCH_ON(aux, "key_scan_kernel", key_scan_kernel, w.nblk, 256, w.nkeys, w.hist, w.start, w.btot, w.bprefix,
                     w.counter, b);
        KS_LAUNCH_ON(aux, "rec_scatter_kernel", rec_scatter_kernel, g256, 256, m, (const int2*)w.keyrank[b], (const int*)w.start,
                     (const int*)w.bprefix, x + c0, y + c0, z + c0, w.rec4[b], w.perm[b]);
        BFE_CUDA(cudaEventRecord(fp->ev_k[b], aux));
        return BFE_OK;
    };
    rc = sort_chunk(0);
    if (rc != BFE_OK) return rc;
    for (int64_t c = 0; c < nchunk; ++c) {
        const int64_t c0 = c * chunk, m = (n - c0) < chunk ? (n - c0) : chunk;
        const int b = (int)(c & 1);
        if (c + 1 < nchunk) { rc = sort_chunk(c + 1); if (rc != BFE_OK) return rc; }
        BFE_CUDA(cudaStreamWaitEvent(stream, fp->ev_k[b], 0));
        const int geval = grid_cap(m, 128, he->num_sms * BFE_PERM_MINB);
        if (staged) {
#define FIELD_STAGE(L, C)                                                                                                        \
    do {                                                                                                                          \
        cudaError_t _e = bfe_launch((field_stage_kernel<6, L, C>), dim3(geval), dim3(128), (size_t)BFE_STAGE_SMEM, stream, nullptr, 0, \
                                    he->g, (const double2*)G4, hs->g, (const double2*)A3, (const double*)hs->xi,                 \
                                    (const double*)hs->p0, facp, m, (const double*)w.rec4[b], crot, srot, w.slot8[b],            \
                                    w.counter + 1 + b);                                                                           \
        if (_e != cudaSuccess) { bfe_set_cuda_error(_e, "field_stage_kernel"); return BFE_ERR_CUDA; }                             \
        bfe_count_launch(1);                                                                                                      \
    } while (0)
            if (hs->g.lmax == 4) { if (cyl) FIELD_STAGE(4, true); else FIELD_STAGE(4, false); }
            else                 { if (cyl) FIELD_STAGE(6, true); else FIELD_STAGE(6, false); }
#undef FIELD_STAGE
        } else {
        // static mode: four tiles per CTA when that still gives >= 8 CTAs per resident slot, else one
        const int64_t ntile_h = (m + 127) / 128;
        // only for the 2^22-point chunks of large sets: at 2^20-point (L2-resident) chunks the persistent grid with tickets is faster
        // (n = 4 x 10^6: 243 vs 261 us per 10^6; n = 10^6: 0.281 vs 0.306 ms)
        const bool st_on = g_bfe_field_eval_static && !(g_bfe_field_support_slim != 0 && nchunk >= 3) && chunk >= ((int64_t)1 << 22);
        const int st_per = st_on ? ((ntile_h >= 8 * 4 * (int64_t)he->num_sms * BFE_PERM_MINB) ? 4 : 1) : 0;
        const int st_grid = st_per ? (int)((ntile_h + st_per - 1) / st_per) : 0;
#define FIELD_REC(L, C, F) KS_LAUNCH("field_rec_kernel", (field_rec_kernel<6, L, C, F>), (st_grid ? st_grid : geval), 128, he->g, G4, hs->g, A3, \
                                     (const double*)hs->xi, (const double*)hs->p0, facp, m, (const double*)w.rec4[b], crot, srot, \
                                     w.slot8[b], w.counter + 1 + b, st_per)
#define FIELD_REC2(L, C) do { if (f32) FIELD_REC(L, C, true); else FIELD_REC(L, C, false); } while (0)
        if (hs->g.lmax == 4) { if (cyl) FIELD_REC2(4, true); else FIELD_REC2(4, false); }
        else                 { if (cyl) FIELD_REC2(6, true); else FIELD_REC2(6, false); }
#undef FIELD_REC2
#undef FIELD_REC
        }
        BFE_CUDA(cudaEventRecord(fp->ev_e[b], stream));
        BFE_CUDA(cudaStreamWaitEvent(gst, fp->ev_e[b], 0));
        const bool slim = g_bfe_field_support_slim != 0 && nchunk >= 3;
        const int sc = g_bfe_field_support_slim > 1 ? g_bfe_field_support_slim : 1;
        const int ggat = slim ? grid_cap(m, 128, he->num_sms * sc) : grid_cap(m, 256, he->num_sms * 4);
        KS_LAUNCH_ON(gst, "field_gather_kernel", field_gather_kernel, ggat, (slim ? 128 : 256), m, n, (const int*)w.perm[b],
                     (const double*)w.slot8[b], out8 + c0);
        if (gsep) BFE_CUDA(cudaEventRecord(fp->ev_g[b], gst));
    }
    BFE_CUDA(cudaEventRecord(fp->ev_out, aux));
    BFE_CUDA(cudaStreamWaitEvent(stream, fp->ev_out, 0));
    if (gsep) {
        BFE_CUDA(cudaEventRecord(fp->ev_out2, gst));
        BFE_CUDA(cudaStreamWaitEvent(stream, fp->ev_out2, 0));
    }
    bfe_kt_end(kt, stream);
    return BFE_OK;
}

// key-ordered integration; the caller has checked eligibility (contracted handles, mmax <= 6, lmax 4 or 6, no trajectory
// output, no apocentre counting)
int bfe_leapfrog_sorted(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, double dt, const double* dt_orbit,
                        double rotfreq, double* state6, int32_t* nsteps_out, cudaStream_t stream) {
    KeySortWs w;
    int rc = keysort_ws(he, hs, stream, norbit, false, w);
    if (rc != BFE_OK) return rc;
    if (norbit > he->orbit_rec_cap || !he->orbit_rec) {
        if (he->orbit_rec) { BFE_CUDA(cudaDeviceSynchronize()); BFE_CUDA(cudaFree(he->orbit_rec)); he->orbit_rec = nullptr; he->orbit_rec_cap = 0; }
        const int64_t cap = norbit + norbit / 8 + 1024;
        BFE_CUDA(cudaMalloc(&he->orbit_rec, sizeof(OrbRec) * (size_t)cap));
        he->orbit_rec_cap = cap;
    }
    OrbRec* rec = (OrbRec*)he->orbit_rec;
    const bool f32 = bfe_use_fp32(he);
    rc = f32 ? bfe_eof_ensure_g4f(he, stream) : bfe_eof_ensure_g4(he, stream);
    // FP64: the per-lane kernels read the per-interval polynomial blocks A4, the TMA-staged ones the three-node blocks A3
    const bool staged = !f32 && g_bfe_stage_eval && he->g.mmax <= 6;
    if (rc == BFE_OK) rc = f32 ? bfe_sl_ensure_a3f(hs, stream) : (staged ? bfe_sl_ensure_a3(hs, stream) : bfe_sl_ensure_a4(hs, stream));
    if (rc != BFE_OK) return rc;
    const void* G4 = f32 ? (const void*)he->g4f : (const void*)he->g4;
    const void* A3 = f32 ? (const void*)hs->a3f : (staged ? (const void*)hs->a3 : (const void*)hs->a4);
    const int g256 = grid_cap(norbit, 256, he->num_sms * 8);
    const int glf = grid_cap(norbit, 128, he->num_sms * BFE_PERM_MINB);
    const int K = g_bfe_orbit_resort > 0 ? g_bfe_orbit_resort : 4;
    const SlFacP facp = bfe_sl_facp(hs);
    const int kt = bfe_kt_begin("leapfrog_sorted_pass", stream);

    KS_LAUNCH("orbit_pack_key_kernel", orbit_pack_key_kernel, g256, 256, he->g, hs->g, w.kc, w.ncell, w.subbits, norbit,
              (const double*)state6, dt, dt_orbit, rec, w.hist, w.keyrank[0]);
    for (int64_t step0 = 0; step0 < nint - 1; step0 += K) {
        const int64_t left = nint - 1 - step0;
        const int k = (int)(left < K ? left : K);
        const int last = (step0 + k >= nint - 1) ? 1 : 0;
        KS_LAUNCH("key_scan_kernel", key_scan_kernel, w.nblk, 256, w.nkeys, w.hist, w.start, w.btot, w.bprefix,
                  w.counter, 0);
        KS_LAUNCH("perm_scatter_kernel", perm_scatter_kernel, g256, 256, norbit, (const int2*)w.keyrank[0],
                  (const int*)w.start, (const int*)w.bprefix, w.perm[0]);
#define LEAP_PERM(L, F) KS_LAUNCH("leapfrog_perm_kernel", (leapfrog_perm_kernel<6, L, F>), glf, 128, he->g, G4, hs->g, A3,     \
                                  (const double*)hs->xi, (const double*)hs->p0, facp, norbit, step0, k,                        \
                                  rotfreq, (int)(step0 == 0), (int)!last, w.kc, w.ncell, w.subbits, (const int*)w.perm[0], rec, w.hist, \
                                  w.keyrank[0], w.counter + 1)
#define LEAP_STAGE(L)                                                                                                            \
    do {                                                                                                                          \
        cudaError_t _e = bfe_launch((leapfrog_stage_kernel<6, L>), dim3(glf), dim3(128), (size_t)BFE_STAGE_SMEM, stream, nullptr, 0, \
                                    he->g, (const double2*)G4, hs->g, (const double2*)A3, (const double*)hs->xi,                 \
                                    (const double*)hs->p0, facp, norbit, step0, k, rotfreq, (int)(step0 == 0), (int)!last,       \
                                    w.kc, w.ncell, w.subbits, (const int*)w.perm[0], rec, w.hist, w.keyrank[0], w.counter + 1);   \
        if (_e != cudaSuccess) { bfe_set_cuda_error(_e, "leapfrog_stage_kernel"); return BFE_ERR_CUDA; }                          \
        bfe_count_launch(1);                                                                                                      \
    } while (0)
        if (staged) { if (hs->g.lmax == 4) LEAP_STAGE(4); else LEAP_STAGE(6); }
        else if (hs->g.lmax == 4) { if (f32) LEAP_PERM(4, true); else LEAP_PERM(4, false); }
        else                      { if (f32) LEAP_PERM(6, true); else LEAP_PERM(6, false); }
#undef LEAP_STAGE
#undef LEAP_PERM
    }
    KS_LAUNCH("orbit_unpack_kernel", orbit_unpack_kernel, g256, 256, norbit, (const OrbRec*)rec, state6, (int*)nsteps_out,
              (int)nint);
    bfe_kt_end(kt, stream);
    return BFE_OK;
}
