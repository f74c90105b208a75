// bfe_orbit_sort.cu -- leapfrog integration of large orbit batches with the orbits kept CELL-COHERENT.
//
// The per-point field evaluation is bound by the L1 tag stage: one cycle per distinct 128-byte line per load
// instruction (ncu l1tex 92 %, profiles/r01_ncu_full_blk_kernels.csv).  When the 32 lanes of a warp sit in the same
// (R, z) table cell they read the same per-cell block, and every load instruction touches ONE line instead of 32
// (the SL interval blocks of such a warp span ~1/2 of the lines): measured 360 -> 156 us per 10^6 points with the
// points merely re-ordered by cell, results bit-identical (profiles/sorted_points_probe.py).  Orbits drift out of
// their cells within a few steps (the vertical cells are thin), so the batch is re-sorted every K steps (option
// "orbit_resort", default 16; measured 0.343 -> 0.288 ns per orbit-step at 10^6 orbits, profiles/orbit_sort_probe.py):
//
//   orbit_pack_kernel      : caller's SoA state (+ step size) -> 64-byte records {x,y,z,vx,vy,vz,dt,index}
//   orbit_cell_hist_kernel : table cell of every orbit -> histogram, last CTA scans (same scheme as bfe_sort.cu)
//   orbit_scatter_kernel   : records to their sorted slots (integer slot claims; two full-sector stores each)
//   leapfrog_rec_kernel    : K velocity-Verlet steps on the sorted records, state in registers; the acceleration at
//                            the chunk's first step is re-evaluated from the position (same inputs, same bits)
//   orbit_unsort_kernel    : records -> caller's SoA order
//
// Used for batches without trajectory output and without apocentre counting (bfe_leapfrog / bfe_leapfrog_dt decide);
// the arithmetic per step is bfe_field_cart_blk, so end states equal the unsorted kernel's bit for bit.
#include "bfe_sortcore.cuh"

struct __align__(32) OrbRec {
    double x, y, z, vx, vy, vz, dt;
    unsigned long long idx;
};

__global__ void __launch_bounds__(256)
orbit_pack_kernel(int64_t n, const double* __restrict__ state6, double dt, const double* __restrict__ dt_orbit,
                  OrbRec* __restrict__ rec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        char* dst = reinterpret_cast<char*>(rec + i);
        bfe_st256(dst, state6[i], state6[n + i], state6[2 * n + i], state6[3 * n + i]);
        bfe_st256(dst + 32, state6[4 * n + i], state6[5 * n + i], dt_orbit ? dt_orbit[i] : dt,
                  __longlong_as_double((long long)i));
    }
}

__global__ void __launch_bounds__(1024)
orbit_cell_hist_kernel(EofGeom g, int ncell, int64_t n, const OrbRec* __restrict__ rec, int* __restrict__ hist,
                       int* __restrict__ cell_start, int* __restrict__ cursor, unsigned int* __restrict__ counter,
                       int* __restrict__ cellid) {
    extern __shared__ int s_hist[];
    __shared__ int s_wsum[32];
    __shared__ bool s_last;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) s_hist[c] = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x)
        cursor[(size_t)c * BFE_CURSOR_STRIDE] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a, b, c, d;
        bfe_ld256_nc(rec + i, a, b, c, d);                   // x, y, z, vx
        int cell;
        bfe_eof_cell_fast(g, a, b, c, cell);                  // an ordering key only: the clamped FP32 index will do
        if (cell < 0 || cell >= ncell) cell = 0;
        atomicAdd(&s_hist[cell], 1);
        cellid[i] = cell;
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
        bfe_block_scan_cells(ncell, hist, cell_start, s_hist, s_wsum);
        if (threadIdx.x == 0) *counter = 0u;
    }
}

__global__ void __launch_bounds__(256)
orbit_scatter_kernel(int64_t n, const OrbRec* __restrict__ src, const int* __restrict__ cellid,
                     const int* __restrict__ cell_start, int* __restrict__ cursor, OrbRec* __restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a0, a1, a2, a3, b0, b1, b2, b3;
        bfe_ld256_nc(src + i, a0, a1, a2, a3);
        bfe_ld256_nc(reinterpret_cast<const char*>(src + i) + 32, b0, b1, b2, b3);
        const int cell = __ldg(cellid + i);
        const int pos = __ldg(cell_start + cell) + atomicAdd(&cursor[(size_t)cell * BFE_CURSOR_STRIDE], 1);   // integer slot claim
        char* d = reinterpret_cast<char*>(dst + pos);
        bfe_st256(d, a0, a1, a2, a3);
        bfe_st256(d + 32, b0, b1, b2, b3);
    }
}

template <int MCAP, int LCAP, bool F32>
__global__ void __launch_bounds__(128)
leapfrog_rec_kernel(EofGeom ge, const void* __restrict__ G4, SlGeom gs, const void* __restrict__ A3,
                    const double* __restrict__ xi, const double* __restrict__ p0tab, const double* __restrict__ fac,
                    int64_t norbit, int64_t step0, int nsteps, double rotfreq, OrbRec* __restrict__ rec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= norbit) return;
    double px, py, pz, vx, vy, vz, dt, idxbits;
    // plain (coherent) loads: this kernel rewrites the record in place
    {
        const double4* r4 = reinterpret_cast<const double4*>(rec + i);
        const double4 a = r4[0], b = r4[1];
        px = a.x; py = a.y; pz = a.z; vx = a.w; vy = b.x; vz = b.y; dt = b.z; idxbits = b.w;
    }
    const double w = BFE_TWOPI * rotfreq;                    // barpos = 2 pi rotfreq (k dt), integrate.py:94-97
    const double hdt2 = 0.5 * (dt * dt);
    double srot, crot;
    sincos(w * ((double)step0 * dt), &srot, &crot);
    CartForce f = bfe_field_cart_blk<MCAP, LCAP, false, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
    double ax = f.fxd + f.fxh, ay = f.fyd + f.fyh, az = f.fzd + f.fzh;
    for (int k = 1; k <= nsteps; ++k) {
        const int64_t step = step0 + k;
        px = px + (vx * dt) + (ax * hdt2);                   // integrate.py:129-131
        py = py + (vy * dt) + (ay * hdt2);
        pz = pz + (vz * dt) + (az * hdt2);
        sincos(w * ((double)step * dt), &srot, &crot);
        f = bfe_field_cart_blk<MCAP, LCAP, false, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
        const double bx = f.fxd + f.fxh, by = f.fyd + f.fyh, bz = f.fzd + f.fzh;             // 134-138
        vx = vx + (0.5 * (ax + bx) * dt);                    // 141-143
        vy = vy + (0.5 * (ay + by) * dt);
        vz = vz + (0.5 * (az + bz) * dt);
        ax = bx; ay = by; az = bz;
    }
    char* d = reinterpret_cast<char*>(rec + i);
    bfe_st256(d, px, py, pz, vx);
    bfe_st256(d + 32, vy, vz, dt, idxbits);
}

__global__ void __launch_bounds__(256)
orbit_unsort_kernel(int64_t n, const OrbRec* __restrict__ rec, double* __restrict__ state6, int* __restrict__ nsteps_out,
                    int nint) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a0, a1, a2, a3, b0, b1, b2, b3;
        bfe_ld256_nc(rec + i, a0, a1, a2, a3);
        bfe_ld256_nc(reinterpret_cast<const char*>(rec + i) + 32, b0, b1, b2, b3);
        const int64_t o = (int64_t)__double_as_longlong(b3);
        state6[o] = a0; state6[n + o] = a1; state6[2 * n + o] = a2;
        state6[3 * n + o] = a3; state6[4 * n + o] = b0; state6[5 * n + o] = b1;
        if (nsteps_out) nsteps_out[o] = nint;
    }
}

int g_bfe_orbit_resort = 16;            // option "orbit_resort": steps between re-sorts (0: never use the sorted path)
int g_bfe_orbit_sort_min = 400000;      // option "orbit_sort_min": smallest batch that takes the sorted path.  The gain needs >~ 32 orbits per
                                        // occupied table cell: 10^6 orbits 3.41 -> 2.93 s (C4), 1.25e5 orbits neutral (0.467 -> 0.463 s on 8 GPUs),
                                        // and below that the fixed cost of a re-sort (~50 us per 16 steps) would show

static size_t os_align(size_t v) { return (v + 255) / 256 * 256; }

// sorted-path integration; the caller has checked eligibility (contracted handles, mmax <= 6, lmax 4 or 6)
int bfe_leapfrog_sorted(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, double dt, const double* dt_orbit,
                        double rotfreq, double* state6, int32_t* nsteps_out, cudaStream_t stream) {
    const int ncell = he->g.numx * he->g.numy;
    const size_t o_hist = 0;
    const size_t o_start = os_align(o_hist + sizeof(int) * ncell);
    const size_t o_cur = os_align(o_start + sizeof(int) * (ncell + 1));
    const size_t o_cid = os_align(o_cur + sizeof(int) * (size_t)ncell * BFE_CURSOR_STRIDE);
    if (norbit > he->orbit_cap || !he->orbit_ws) {
        if (he->orbit_ws) { BFE_CUDA(cudaDeviceSynchronize()); BFE_CUDA(cudaFree(he->orbit_ws)); he->orbit_ws = nullptr; }
        const int64_t cap = norbit + norbit / 8 + 1024;
        const size_t o_a = os_align(o_cid + sizeof(int) * (size_t)cap);
        BFE_CUDA(cudaMalloc(&he->orbit_ws, o_a + 2 * sizeof(OrbRec) * (size_t)cap));
        BFE_CUDA(cudaMemset(he->orbit_ws, 0, o_cid));
        BFE_CUDA(cudaDeviceSynchronize());
        he->orbit_cap = cap;
    }
    char* b = (char*)he->orbit_ws;
    int* hist = (int*)(b + o_hist); int* cell_start = (int*)(b + o_start); int* cursor = (int*)(b + o_cur);
    int* cellid = (int*)(b + o_cid);
    const size_t o_a = os_align(o_cid + sizeof(int) * (size_t)he->orbit_cap);
    OrbRec* bufs[2] = {(OrbRec*)(b + o_a), (OrbRec*)(b + o_a) + he->orbit_cap};

    const bool f32 = g_bfe_table_fp32 != 0;
    int rc = f32 ? bfe_eof_ensure_g4f(he, stream) : bfe_eof_ensure_g4(he, stream);
    if (rc == BFE_OK) rc = f32 ? bfe_sl_ensure_a3f(hs, stream) : bfe_sl_ensure_a3(hs, stream);
    if (rc != BFE_OK) return rc;
    const void* G4 = f32 ? (const void*)he->g4f : (const void*)he->g4;
    const void* A3 = f32 ? (const void*)hs->a3f : (const void*)hs->a3;

    const int per = (ncell + 1023) / 1024;
    const size_t ss = sizeof(int) * (size_t)per * 1024;
    if (ss > 200 * 1024) return BFE_ERR_UNSUPPORTED;
    if (ss > 48 * 1024)
        BFE_CUDA(cudaFuncSetAttribute(orbit_cell_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss));
    int ghist = (int)((norbit + 2047) / 2048);
    if (ghist > he->num_sms) ghist = he->num_sms;
    if (ghist < 1) ghist = 1;
    int g256 = (int)((norbit + 255) / 256);
    if (g256 > he->num_sms * 8) g256 = he->num_sms * 8;
    const int glf = (int)((norbit + 127) / 128);

    orbit_pack_kernel<<<g256, 256, 0, stream>>>(norbit, state6, dt, dt_orbit, bufs[0]);
    BFE_LAUNCH_CHECK("orbit_pack_kernel");
    int cur = 0;
    const int K = g_bfe_orbit_resort > 0 ? g_bfe_orbit_resort : 32;
    for (int64_t step0 = 0; step0 < nint - 1; step0 += K) {
        const int k = (int)((nint - 1 - step0) < K ? (nint - 1 - step0) : K);
        orbit_cell_hist_kernel<<<ghist, 1024, ss, stream>>>(he->g, ncell, norbit, bufs[cur], hist, cell_start, cursor,
                                                           he->counter, cellid);
        BFE_LAUNCH_CHECK("orbit_cell_hist_kernel");
        orbit_scatter_kernel<<<g256, 256, 0, stream>>>(norbit, bufs[cur], cellid, cell_start, cursor, bufs[cur ^ 1]);
        BFE_LAUNCH_CHECK("orbit_scatter_kernel");
        cur ^= 1;
#define LEAP_REC(L, F) leapfrog_rec_kernel<6, L, F><<<glf, 128, 0, stream>>>(he->g, G4, hs->g, A3, hs->xi, hs->p0, hs->fac, norbit, \
                                                                         step0, k, rotfreq, bufs[cur])
        if (hs->g.lmax == 4) { if (f32) LEAP_REC(4, true); else LEAP_REC(4, false); }
        else                 { if (f32) LEAP_REC(6, true); else LEAP_REC(6, false); }
#undef LEAP_REC
        BFE_LAUNCH_CHECK("leapfrog_rec_kernel");
    }
    orbit_unsort_kernel<<<g256, 256, 0, stream>>>(norbit, bufs[cur], state6, nsteps_out, (int)nint);
    BFE_LAUNCH_CHECK("orbit_unsort_kernel");
    return BFE_OK;
}


// ---------------------------------------------------------------------------
// One-shot evaluation of many points (Fields.return_forces_cart / _cyl) in cell order:
//   point_cell_hist_kernel  : cell of every point -> histogram (+ scan by the last CTA), cell id kept
//   point_scatter_kernel    : 32-byte records {x, y, z, index} to their sorted slots, inverse permutation
//   field_rec_kernel        : the field at the sorted records -> 64-byte result slots in sorted order (coalesced)
//   field_gather_kernel     : caller's order: slot through the inverse permutation -> the eight SoA outputs (coalesced)
// in chunks of g_bfe_field_sort_chunk points (bounded workspace).  Per 10^6 points: sort ~45 us + evaluation ~155 us + gather
// ~30 us against 358 us for the evaluation in the caller's order -- for DISC-like point sets; see g_bfe_field_sort_min.
// ---------------------------------------------------------------------------
int g_bfe_field_sort_chunk = 4 << 20;   // option "field_sort_chunk": points per sort + evaluate + gather pass

__global__ void __launch_bounds__(1024)
point_cell_hist_kernel(EofGeom g, int ncell, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                       const double* __restrict__ z, int* __restrict__ hist, int* __restrict__ cell_start,
                       int* __restrict__ cursor, unsigned int* __restrict__ counter, int* __restrict__ cellid) {
    extern __shared__ int s_hist[];
    __shared__ int s_wsum[32];
    __shared__ bool s_last;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) s_hist[c] = 0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x)
        cursor[(size_t)c * BFE_CURSOR_STRIDE] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int cell;
        bfe_eof_cell_fast(g, __ldg(x + i), __ldg(y + i), __ldg(z + i), cell);     // ordering key only
        if (cell < 0 || cell >= ncell) cell = 0;
        atomicAdd(&s_hist[cell], 1);
        cellid[i] = cell;
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
        bfe_block_scan_cells(ncell, hist, cell_start, s_hist, s_wsum);
        if (threadIdx.x == 0) *counter = 0u;
    }
}

__global__ void __launch_bounds__(256)
point_scatter_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                     const int* __restrict__ cellid, const int* __restrict__ cell_start, int* __restrict__ cursor,
                     double* __restrict__ rec4, int* __restrict__ inv) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        const int cell = __ldg(cellid + i);
        const int pos = __ldg(cell_start + cell) + atomicAdd(&cursor[(size_t)cell * BFE_CURSOR_STRIDE], 1);   // integer slot claim
        bfe_st256(rec4 + 4 * (size_t)pos, px, py, pz, __longlong_as_double((long long)i));
        inv[i] = pos;
    }
}

template <int MCAP, int LCAP, bool CYL, bool F32>
__global__ void __launch_bounds__(128)
field_rec_kernel(EofGeom ge, const void* __restrict__ G4, SlGeom gs, const void* __restrict__ A3,
                 const double* __restrict__ xi, const double* __restrict__ p0tab, const double* __restrict__ fac,
                 int64_t n, const double* __restrict__ rec4, double crot, double srot, double* __restrict__ slot8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px, py, pz, id_;
        bfe_ld256_nc(rec4 + 4 * (size_t)i, px, py, pz, id_);
        const CartForce f = bfe_field_cart_blk<MCAP, LCAP, CYL, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
        char* d = reinterpret_cast<char*>(slot8 + 8 * (size_t)i);
        bfe_st256(d, f.fxd, f.fxh, f.fyd, f.fyh);
        bfe_st256(d + 32, f.fzd, f.fzh, f.pd, f.ph);
    }
}

__global__ void __launch_bounds__(256)
field_gather_kernel(int64_t n, int64_t ntot, const int* __restrict__ inv, const double* __restrict__ slot8,
                    double* __restrict__ out8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const char* s = reinterpret_cast<const char*>(slot8 + 8 * (size_t)__ldg(inv + i));
        double a0, a1, a2, a3, b0, b1, b2, b3;
        bfe_ld256_nc(s, a0, a1, a2, a3);
        bfe_ld256_nc(s + 32, b0, b1, b2, b3);
        out8[i] = a0; out8[ntot + i] = a1; out8[2 * ntot + i] = a2; out8[3 * ntot + i] = a3;
        out8[4 * ntot + i] = b0; out8[5 * ntot + i] = b1; out8[6 * ntot + i] = b2; out8[7 * ntot + i] = b3;
    }
}

int g_bfe_field_sort_min = 0;           // option "field_sort_min": smallest point set evaluated in cell order; 0 (default): never --
                                        // measured per 10^6 points (profiles/field_sort_probe.py): disc points 358 -> 284 us, but halo points
                                        // 349 -> 437 us (hot edge cells serialise the slot claims, no SL coherence inside a cell) and no gain
                                        // with FP32 tables: the sort + gather (~90 us) only pays for disc-like sets

int bfe_field_force_sorted(bfe_eof* he, bfe_sl* hs, int64_t n, const double* x, const double* y, const double* z,
                           double crot, double srot, double* out8, bool cyl, cudaStream_t stream) {
    const int ncell = he->g.numx * he->g.numy;
    const int64_t chunk_opt = g_bfe_field_sort_chunk > 0 ? g_bfe_field_sort_chunk : (4 << 20);
    const int64_t chunk = n < chunk_opt ? n : chunk_opt;
    const size_t o_hist = 0;
    const size_t o_start = os_align(o_hist + sizeof(int) * ncell);
    const size_t o_cur = os_align(o_start + sizeof(int) * (ncell + 1));
    const size_t o_cid = os_align(o_cur + sizeof(int) * (size_t)ncell * BFE_CURSOR_STRIDE);
    // the orbit workspace is shared: header + cell ids + 2 x 64 B per unit covers cell ids + inv + 32-B records + 64-B slots
    if (chunk > he->orbit_cap || !he->orbit_ws) {
        if (he->orbit_ws) { BFE_CUDA(cudaDeviceSynchronize()); BFE_CUDA(cudaFree(he->orbit_ws)); he->orbit_ws = nullptr; }
        const int64_t cap = chunk + chunk / 8 + 1024;
        const size_t o_a = os_align(o_cid + sizeof(int) * (size_t)cap);
        BFE_CUDA(cudaMalloc(&he->orbit_ws, o_a + 2 * sizeof(OrbRec) * (size_t)cap));
        BFE_CUDA(cudaMemset(he->orbit_ws, 0, o_cid));
        BFE_CUDA(cudaDeviceSynchronize());
        he->orbit_cap = cap;
    }
    char* b = (char*)he->orbit_ws;
    int* hist = (int*)(b + o_hist); int* cell_start = (int*)(b + o_start); int* cursor = (int*)(b + o_cur);
    int* cellid = (int*)(b + o_cid);
    const size_t o_a = os_align(o_cid + sizeof(int) * (size_t)he->orbit_cap);
    double* slot8 = (double*)(b + o_a);                                          // 64 B per point
    double* rec4 = (double*)(b + o_a + sizeof(OrbRec) * (size_t)he->orbit_cap);  // 32 B per point
    int* inv = (int*)(b + o_a + sizeof(OrbRec) * (size_t)he->orbit_cap + 32 * (size_t)he->orbit_cap);   // 4 B per point

    const bool f32 = g_bfe_table_fp32 != 0;
    int rc = f32 ? bfe_eof_ensure_g4f(he, stream) : bfe_eof_ensure_g4(he, stream);
    if (rc == BFE_OK) rc = f32 ? bfe_sl_ensure_a3f(hs, stream) : bfe_sl_ensure_a3(hs, stream);
    if (rc != BFE_OK) return rc;
    const void* G4 = f32 ? (const void*)he->g4f : (const void*)he->g4;
    const void* A3 = f32 ? (const void*)hs->a3f : (const void*)hs->a3;
    const int per = (ncell + 1023) / 1024;
    const size_t ss = sizeof(int) * (size_t)per * 1024;
    if (ss > 200 * 1024) return BFE_ERR_UNSUPPORTED;
    if (ss > 48 * 1024)
        BFE_CUDA(cudaFuncSetAttribute(point_cell_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ss));
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        const int64_t m = (n - c0) < chunk ? (n - c0) : chunk;
        int ghist = (int)((m + 2047) / 2048);
        if (ghist > he->num_sms) ghist = he->num_sms;
        if (ghist < 1) ghist = 1;
        int g256 = (int)((m + 255) / 256);
        if (g256 > he->num_sms * 8) g256 = he->num_sms * 8;
        int64_t need = (m + 127) / 128, cap = (int64_t)he->num_sms * 16;
        const int geval = (int)(need < cap ? need : cap);
        point_cell_hist_kernel<<<ghist, 1024, ss, stream>>>(he->g, ncell, m, x + c0, y + c0, z + c0, hist, cell_start, cursor,
                                                           he->counter, cellid);
        BFE_LAUNCH_CHECK("point_cell_hist_kernel");
        point_scatter_kernel<<<g256, 256, 0, stream>>>(m, x + c0, y + c0, z + c0, cellid, cell_start, cursor, rec4, inv);
        BFE_LAUNCH_CHECK("point_scatter_kernel");
#define FIELD_REC(L, C, F) field_rec_kernel<6, L, C, F><<<geval, 128, 0, stream>>>(he->g, G4, hs->g, A3, hs->xi, hs->p0, hs->fac, m, \
                                                                                rec4, crot, srot, slot8)
#define FIELD_REC2(L, C) do { if (f32) FIELD_REC(L, C, true); else FIELD_REC(L, C, false); } while (0)
        if (hs->g.lmax == 4) { if (cyl) FIELD_REC2(4, true); else FIELD_REC2(4, false); }
        else                 { if (cyl) FIELD_REC2(6, true); else FIELD_REC2(6, false); }
#undef FIELD_REC2
#undef FIELD_REC
        BFE_LAUNCH_CHECK("field_rec_kernel");
        field_gather_kernel<<<g256, 256, 0, stream>>>(m, n, inv, slot8, out8 + c0);
        BFE_LAUNCH_CHECK("field_gather_kernel");
    }
    return BFE_OK;
}
