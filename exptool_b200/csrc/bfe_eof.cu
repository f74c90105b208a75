// bfe_eof.cu -- EOF (cylindrical disc basis) kernels.
//
//   eof_relayout_acc_kernel : [m][n][node] tables -> node-major accumulate table
//   eof_accumulate_kernel   : eof.accumulate (eof.py:492-551), direct formulation
//   eof_contract_kernel     : G[node][m][6] = sum_n coef[m,n] T_f[m,n,node]
//   eof_force_kernel        : eof.accumulated_eval_particles (eof.py:989-1144)
//   eof_points_kernel       : eof.force_eval (eof.py:756-870)
#include "bfe_device.cuh"

// ---------------------------------------------------------------------------
// table re-layout: T_acc[node][j], j < nch:  j = m*norder+n (cos, from potC),
// j = ncos + (m-1)*norder + n (sin, m >= 1, from potS); padding channels are zero.
// ---------------------------------------------------------------------------
__global__ void eof_relayout_acc_kernel(EofGeom g, const double* __restrict__ potC, const double* __restrict__ potS,
                                        int nch, int nch_pad, double* __restrict__ t_acc) {
    // block: 32 nodes x 32 channels tile through shared memory (coalesced both sides)
    __shared__ double tile[32][33];
    const int node0 = blockIdx.x * 32, ch0 = blockIdx.y * 32;
    const int ncos = (g.mmax + 1) * g.norder;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int ch = ch0 + r, node = node0 + threadIdx.x;
        double v = 0.0;
        if (ch < nch && node < g.nnode) {
            if (ch < ncos) v = potC[(size_t)ch * g.nnode + node];
            else           v = potS[(size_t)(ch - ncos + g.norder) * g.nnode + node];
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int node = node0 + r, ch = ch0 + threadIdx.x;
        if (node < g.nnode && ch < nch_pad) t_acc[(size_t)node * nch_pad + ch] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------------------
// accumulate, direct formulation.
//
// One thread owns one (m,n,trig) channel for the whole launch and keeps its partial
// sum in a register; a CTA walks tiles of TILE particles.  Phase A (thread per particle)
// computes the bin, the four mass-weighted bilinear weights and mass-free cos/sin(m phi)
// into shared memory; phase B (thread per channel) reads, for every particle of the tile,
// the four corner rows of the node-major table -- each row is nch contiguous doubles, so a
// warp's loads are fully coalesced 256-B segments -- and does 5 DFMA.  No atomics and no
// shuffles in the loop; per-CTA partials are combined by the last CTA to finish.
// ---------------------------------------------------------------------------
template <int MCAP, int TILE, int KCH>
__global__ void __launch_bounds__(256)
eof_accumulate_kernel(EofGeom g, const double* __restrict__ t_acc, int nch, int nch_pad,
                      int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                      const double* __restrict__ z, const double* __restrict__ mass,
                      double* __restrict__ partial, unsigned int* __restrict__ counter,
                      double* __restrict__ cos_out, double* __restrict__ sin_out) {
    constexpr int NTRIG = 2 * MCAP + 1;
    __shared__ int s_node[TILE];
    __shared__ double s_w[4][TILE];
    __shared__ double s_trig[NTRIG][TILE + 1];  // [0..mmax] = cos(m phi), [mmax+1..2mmax] = sin(m phi), m>=1;
                                                // +1: rows read by one warp at the same p hit different banks
    __shared__ bool s_last;

    const int tid = threadIdx.x;
    const int ncos = (g.mmax + 1) * g.norder;
    // channels owned by this thread: tid + k*256
    int trig_idx[KCH];
    double acc[KCH];
#pragma unroll
    for (int k = 0; k < KCH; ++k) {
        int ch = tid + k * 256;
        trig_idx[k] = 0;
        if (ch < nch) trig_idx[k] = (ch < ncos) ? ch / g.norder : g.mmax + 1 + (ch - ncos) / g.norder;
        acc[k] = 0.0;
    }
    const double* __restrict__ tcol = t_acc + tid;
    const int rowstep = nch_pad;                 // iy+1
    const int colstep = g.ny1 * nch_pad;         // ix+1

    const int64_t ntiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- phase A
        if (tid < TILE) {
            int64_t ip = tile * TILE + tid;
            double px = 0.0, py = 0.0, pz = 0.0, pm = 0.0;
            if (ip < n) { px = __ldg(x + ip); py = __ldg(y + ip); pz = __ldg(z + ip); pm = __ldg(mass + ip); }
            double r = sqrt(px * px + py * py + 1.e-10);          // eof.py:531
            EofBin b = bfe_eof_bin(g, r, pz);
            double c1, s1;
            bfe_cossin_phi(px, py, c1, s1);                       // eof.py:532
            s_node[tid] = b.node;
            s_w[0][tid] = b.c00 * pm;                             // mass folded into the weights (eof.py:536)
            s_w[1][tid] = b.c10 * pm;
            s_w[2][tid] = b.c01 * pm;
            s_w[3][tid] = b.c11 * pm;
            double cm = 1.0, sm = 0.0;
            s_trig[0][tid] = 1.0;
#pragma unroll
            for (int m = 1; m <= MCAP; ++m) {
                if (m <= g.mmax) {
                    double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1;
                    cm = cn; sm = sn;
                    s_trig[m][tid] = cm;
                    s_trig[g.mmax + m][tid] = sm;
                }
            }
        }
        __syncthreads();
        // ---- phase B
#pragma unroll
        for (int k = 0; k < KCH; ++k) {
            if (tid + k * 256 < nch) {
                const double* trow = &s_trig[trig_idx[k]][0];
                const double* tk = tcol + k * 256;
                double a = acc[k];
#pragma unroll 4
                for (int p = 0; p < TILE; ++p) {
                    const double* base = tk + (size_t)s_node[p] * nch_pad;
                    double t00 = __ldg(base);
                    double t01 = __ldg(base + rowstep);
                    double t10 = __ldg(base + colstep);
                    double t11 = __ldg(base + colstep + rowstep);
                    double v = t00 * s_w[0][p] + t10 * s_w[1][p] + t01 * s_w[2][p] + t11 * s_w[3][p];
                    a = fma(trow[p], v, a);
                }
                acc[k] = a;
            }
        }
        __syncthreads();
    }

    // ---- per-CTA partial, last CTA reduces in fixed order (deterministic)
#pragma unroll
    for (int k = 0; k < KCH; ++k)
        if (tid + k * 256 < nch_pad) partial[(size_t)blockIdx.x * nch_pad + tid + k * 256] = acc[k];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        unsigned int done = atomicAdd(counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int ch = tid; ch < nch; ch += 256) {
            double s = bfe_column_sum(partial, (int)gridDim.x, nch_pad, ch) * BFE_FOURPI_NEG;   // eof.py:526,550
            if (ch < ncos) cos_out[ch] = s;
            else           sin_out[ch - ncos + g.norder] = s;
        }
        for (int k = tid; k < g.norder; k += 256) sin_out[k] = 0.0;  // m = 0 sine row (potS[0] == 0, eof.py:293)
        if (tid == 0) *counter = 0u;
    }
}

// ---------------------------------------------------------------------------
// contraction with a coefficient set.
// grid: (ceil(nnode/128), (mmax+1)*6); thread per node, loops n (loads coalesced over nodes).
// q = field*2 + trig : 0 pc, 1 ps, 2 rc, 3 rs, 4 zc, 5 zs.  t_force = [potC,rfC,zfC,potS,rfS,zfS].
// ---------------------------------------------------------------------------
// Two nodes per thread, six terms each per pass: 12 independent table loads in flight per thread, and the grid
// (33 x 42 CTAs for the 129 x 65 table) is resident in ONE wave -- with one node per thread it was 1.17 waves of
// short-lived CTAs, i.e. the time of two (ncu: 36 % of DRAM peak, long-scoreboard 38 cycles per issue).
__global__ void __launch_bounds__(128)
eof_contract_kernel(EofGeom g, const double* __restrict__ t_force, size_t tab_elems,
                    const double* __restrict__ cosc, const double* __restrict__ sinc,
                    int m1, int m2, int nuse, int no_odd,
                    double* __restrict__ G, int gstride, int deep) {
    const int node = blockIdx.x * 256 + threadIdx.x, node2 = node + 128;
    const int m = blockIdx.y / 6, q = blockIdx.y % 6;
    bfe_pdl_wait();
    bfe_pdl_trigger();
    if (node >= g.nnode) return;
    const bool two = node2 < g.nnode;
    const int field = q >> 1, trig = q & 1;
    double s = 0.0, r = 0.0;
    bool use = (m >= m1) && (m <= m2) && !(no_odd && (m & 1)) && !(trig == 1 && m == 0);
    if (use) {
        const double* T = t_force + (size_t)(trig * 3 + field) * tab_elems + (size_t)m * g.norder * g.nnode + node;
        const double* T2 = two ? T + 128 : T;
        const double* c = (trig ? sinc : cosc) + m * g.norder;
        const int nn = nuse < g.norder ? nuse : g.norder;
        double s1 = 0.0, s2 = 0.0, r1 = 0.0, r2 = 0.0;
        int k = 0;
        if (deep) {
            for (; k + 8 < nn; k += 9) {
                double t[9], v[9];
#pragma unroll
                for (int u = 0; u < 9; ++u) { t[u] = __ldg(T + (size_t)(k + u) * g.nnode); v[u] = __ldg(T2 + (size_t)(k + u) * g.nnode); }
#pragma unroll
                for (int u = 0; u < 9; u += 3) {
                    const double c0 = __ldg(c + k + u), c1 = __ldg(c + k + u + 1), c2 = __ldg(c + k + u + 2);
                    s = fma(c0, t[u], s); s1 = fma(c1, t[u + 1], s1); s2 = fma(c2, t[u + 2], s2);
                    r = fma(c0, v[u], r); r1 = fma(c1, v[u + 1], r1); r2 = fma(c2, v[u + 2], r2);
                }
            }
        }
        for (; k + 5 < nn; k += 6) {
            double t[6], v[6], cc[6];
#pragma unroll
            for (int u = 0; u < 6; ++u) { t[u] = __ldg(T + (size_t)(k + u) * g.nnode); v[u] = __ldg(T2 + (size_t)(k + u) * g.nnode); cc[u] = __ldg(c + k + u); }
            s = fma(cc[0], t[0], s); s1 = fma(cc[1], t[1], s1); s2 = fma(cc[2], t[2], s2);
            s = fma(cc[3], t[3], s); s1 = fma(cc[4], t[4], s1); s2 = fma(cc[5], t[5], s2);
            r = fma(cc[0], v[0], r); r1 = fma(cc[1], v[1], r1); r2 = fma(cc[2], v[2], r2);
            r = fma(cc[3], v[3], r); r1 = fma(cc[4], v[4], r1); r2 = fma(cc[5], v[5], r2);
        }
        for (; k < nn; ++k) { const double ck = __ldg(c + k); s = fma(ck, __ldg(T + (size_t)k * g.nnode), s); r = fma(ck, __ldg(T2 + (size_t)k * g.nnode), r); }
        s += s1 + s2;
        r += r1 + r2;
    }
    G[(size_t)node * gstride + m * 6 + q] = s;
    if (two) G[(size_t)node2 * gstride + m * 6 + q] = r;
}

// ---------------------------------------------------------------------------
// field evaluation at particles (thread per particle)
// ---------------------------------------------------------------------------
template <int MCAP>
__global__ void __launch_bounds__(128)
eof_force_kernel(EofGeom g, const double* __restrict__ G, int gstride, int64_t n,
                 const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                 double* __restrict__ p0, double* __restrict__ p, double* __restrict__ fr,
                 double* __restrict__ fp, double* __restrict__ fz, double* __restrict__ R) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        double r = sqrt(px * px + py * py + 1.e-10);              // eof.py:1070
        EofBin b = bfe_eof_bin(g, r, pz);
        double c1, s1;
        bfe_cossin_phi(px, py, c1, s1);                           // eof.py:1068
        EofField f = bfe_eof_eval<MCAP>(g, G, gstride, b, c1, s1);
        p0[i] = f.p0; p[i] = f.p; fr[i] = f.fr; fp[i] = f.fp; fz[i] = f.fz; R[i] = r;
    }
}

// g_con [node][m][6] -> G4 [cell][m][corner][3 double2]
__global__ void eof_expand_g4_kernel(EofGeom g, const double* __restrict__ G, int gstride, double2* __restrict__ G4) {
    const int per = 12 * (g.mmax + 1);
    const int64_t total = (int64_t)g.numx * g.numy * per;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int cell = (int)(t / per), c = (int)(t - (int64_t)cell * per);
        const int m = c / 12, r = c - m * 12, corner = r / 3, part = r - corner * 3;
        const int ix = cell / g.numy, iy = cell - ix * g.numy;
        const int node = ix * g.ny1 + iy + ((corner & 1) ? g.ny1 : 0) + ((corner & 2) ? 1 : 0);   // 00,10,01,11
        const double* src = G + (size_t)node * gstride + m * 6 + part * 2;
        G4[t] = make_double2(src[0], src[1]);
    }
}

// g_con [node][m][6] -> G4f [cell][m][corner][6 floats]  (table_fp32 mode)
__global__ void eof_expand_g4f_kernel(EofGeom g, const double* __restrict__ G, int gstride, float* __restrict__ G4f) {
    const int per = 24 * (g.mmax + 1);
    const int64_t total = (int64_t)g.numx * g.numy * per;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int cell = (int)(t / per), c = (int)(t - (int64_t)cell * per);
        const int m = c / 24, r = c - m * 24, corner = r / 6, part = r - corner * 6;
        const int ix = cell / g.numy, iy = cell - ix * g.numy;
        const int node = ix * g.ny1 + iy + ((corner & 1) ? g.ny1 : 0) + ((corner & 2) ? 1 : 0);   // 00,10,01,11
        G4f[t] = (float)G[(size_t)node * gstride + m * 6 + part];
    }
}

// warp-cooperative versions (bfe_eof_eval_staged): lanes of a warp take 32 consecutive points
template <int MCAP>
__global__ void __launch_bounds__(128)
eof_force_staged_kernel(EofGeom g, const double2* __restrict__ G4, int64_t n,
                        const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                        double* __restrict__ p0, double* __restrict__ p, double* __restrict__ fr,
                        double* __restrict__ fp, double* __restrict__ fz, double* __restrict__ R) {
    __shared__ double2 s_stage[4][BFE_STAGE_DOUBLE2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* st = s_stage[warp];
    const int64_t wglobal = (int64_t)blockIdx.x * 4 + warp, wtotal = (int64_t)gridDim.x * 4;
    for (int64_t base = wglobal * 32; base < n; base += wtotal * 32) {
        const int64_t i = base + lane;
        const bool on = i < n;
        const int64_t ii = on ? i : n - 1;
        double px = __ldg(x + ii), py = __ldg(y + ii), pz = __ldg(z + ii);
        double r = sqrt(px * px + py * py + 1.e-10);              // eof.py:1070
        EofBin b = bfe_eof_bin(g, r, pz);
        double c1, s1;
        bfe_cossin_phi(px, py, c1, s1);
        EofField f = bfe_eof_eval_staged<MCAP>(g, G4, b, c1, s1, st, lane);
        if (on) { p0[i] = f.p0; p[i] = f.p; fr[i] = f.fr; fp[i] = f.fp; fz[i] = f.fz; R[i] = r; }
    }
}

template <int MCAP>
__global__ void __launch_bounds__(128)
eof_points_staged_kernel(EofGeom g, const double2* __restrict__ G4, int64_t n,
                         const double* __restrict__ r, const double* __restrict__ z, const double* __restrict__ phi,
                         double* __restrict__ fr, double* __restrict__ fp, double* __restrict__ fz,
                         double* __restrict__ p, double* __restrict__ p0) {
    __shared__ double2 s_stage[4][BFE_STAGE_DOUBLE2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* st = s_stage[warp];
    const int64_t wglobal = (int64_t)blockIdx.x * 4 + warp, wtotal = (int64_t)gridDim.x * 4;
    for (int64_t base = wglobal * 32; base < n; base += wtotal * 32) {
        const int64_t i = base + lane;
        const bool on = i < n;
        const int64_t ii = on ? i : n - 1;
        EofBin b = bfe_eof_bin(g, __ldg(r + ii), __ldg(z + ii));
        double c1, s1;
        sincos(__ldg(phi + ii), &s1, &c1);
        EofField f = bfe_eof_eval_staged<MCAP>(g, G4, b, c1, s1, st, lane);
        if (on) { fr[i] = f.fr; fp[i] = f.fp; fz[i] = f.fz; p[i] = f.p + f.p0; p0[i] = f.p0; }
    }
}

// eof.force_eval at (r, z, phi) points: returns fr, fp, fz (m=0 included), p+p0, p0
template <int MCAP>
__global__ void __launch_bounds__(128)
eof_points_kernel(EofGeom g, const double* __restrict__ G, int gstride, int64_t n,
                  const double* __restrict__ r, const double* __restrict__ z, const double* __restrict__ phi,
                  double* __restrict__ fr, double* __restrict__ fp, double* __restrict__ fz,
                  double* __restrict__ p, double* __restrict__ p0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        EofBin b = bfe_eof_bin(g, __ldg(r + i), __ldg(z + i));
        double c1, s1;
        sincos(__ldg(phi + i), &s1, &c1);
        EofField f = bfe_eof_eval<MCAP>(g, G, gstride, b, c1, s1);
        fr[i] = f.fr; fp[i] = f.fp; fz[i] = f.fz; p[i] = f.p + f.p0; p0[i] = f.p0;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int grid_for(int64_t n, int block, int num_sms, int per_sm) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)num_sms * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" int bfe_eof_create(const bfe_eof_params* p, const double* potC, const double* rforceC,
                              const double* zforceC, const double* potS, const double* rforceS,
                              const double* zforceS, void* stream_, bfe_eof** out) {
    if (!p || !out || !potC || !potS) return BFE_ERR_ARG;
    if (p->mmax < 0 || p->mmax > BFE_MAX_MMAX || p->norder < 1 || p->numx < 1 || p->numy < 1) return BFE_ERR_ARG;
    if (p->cmap < 0 || p->cmap > 2) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    bfe_eof* h = new bfe_eof();
    h->par = *p;
    EofGeom& g = h->g;
    g.mmax = p->mmax; g.norder = p->norder; g.numx = p->numx; g.numy = p->numy; g.cmap = p->cmap;
    g.ny1 = p->numy + 1;
    g.nnode = (p->numx + 1) * (p->numy + 1);
    g.xmin = p->xmin; g.dx = p->dx; g.ymin = p->ymin; g.dy = p->dy; g.ascale = p->ascale; g.hscale = p->hscale;
    g.inv_dx = 1.0 / p->dx; g.inv_dy = 1.0 / p->dy;
    g.inv_ascale = 1.0 / p->ascale; g.inv_hscale = 1.0 / p->hscale;
    BFE_CUDA(cudaGetDevice(&h->device));
    BFE_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
    h->nch = (2 * p->mmax + 1) * p->norder;
    h->nch_pad = (h->nch + 31) / 32 * 32;
    if (h->nch_pad > 1024) { delete h; return BFE_ERR_UNSUPPORTED; }  // <= 4 channels per thread, 256-thread CTAs
    h->tab_elems = (size_t)(p->mmax + 1) * p->norder * g.nnode;
    h->gstride = 6 * (p->mmax + 1);
    h->contracted = 0;
    h->sort_cap = 0; h->sort_ws = nullptr; h->prepared_n = -1; h->prepared_has_mass = 0;
    h->host_pipe = nullptr;
    h->owns_tables = 1;
    h->max_ctas = h->num_sms * 4;
    BFE_CUDA(cudaMalloc(&h->t_acc, (size_t)g.nnode * h->nch_pad * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->g_con, (size_t)g.nnode * h->gstride * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->g4, (size_t)g.numx * g.numy * 12 * (p->mmax + 1) * 2 * sizeof(double)));
    h->g4_valid = 0;
    h->g4f = nullptr; h->g4f_valid = 0; h->table_fp32 = -1;
    h->orbit_ws = nullptr; h->orbit_cap = 0; h->orbit_hdr = 0; h->keycell = nullptr; h->keycell_nkeys = 0; h->keycell_nkeys2 = 0; h->keycell_tag = 0; h->orbit_rec = nullptr; h->orbit_rec_cap = 0; h->field_pipe = nullptr;
    BFE_CUDA(cudaMalloc(&h->partial, (size_t)(h->max_ctas + 64) * h->nch_pad * sizeof(double)));   // + group rows of the two-level reduce
    BFE_CUDA(cudaMalloc(&h->counter, 128 * sizeof(unsigned int)));      // [0] last-CTA, [1] task queue, [64..] reduce groups
    BFE_CUDA(cudaMemsetAsync(h->counter, 0, 128 * sizeof(unsigned int), stream));
    h->t_force = nullptr;
    if (rforceC && zforceC && rforceS && zforceS) {
        BFE_CUDA(cudaMalloc(&h->t_force, 6 * h->tab_elems * sizeof(double)));
        const double* src[6] = {potC, rforceC, zforceC, potS, rforceS, zforceS};
        for (int k = 0; k < 6; ++k)
            BFE_CUDA(cudaMemcpyAsync(h->t_force + k * h->tab_elems, src[k], h->tab_elems * sizeof(double),
                                     cudaMemcpyDeviceToDevice, stream));
    }
    dim3 blk(32, 8), grd((g.nnode + 31) / 32, h->nch_pad / 32);
    eof_relayout_acc_kernel<<<grd, blk, 0, stream>>>(g, potC, potS, h->nch, h->nch_pad, h->t_acc);
    BFE_LAUNCH_CHECK("eof_relayout_acc_kernel");
    *out = h;
    return BFE_OK;
}

// A second handle on the SAME device tables with its own contraction, workspaces and counters, so that two
// independent particle sets (e.g. consecutive snapshots of a time series) can be in flight on two streams
// and fill each other's launch gaps, tails and latency-bound phases.  The parent must outlive its clones.
extern "C" void bfe_eof_destroy(bfe_eof* h);
extern "C" int bfe_eof_clone(const bfe_eof* src, void* stream_, bfe_eof** out) {
    if (!src || !out) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    bfe_eof* h = new bfe_eof(*src);
    h->owns_tables = 0;
    h->contracted = 0; h->g4_valid = 0; h->g4f = nullptr; h->g4f_valid = 0;
    h->sort_cap = 0; h->sort_ws = nullptr; h->prepared_n = -1; h->prepared_has_mass = 0;
    h->host_pipe = nullptr;
    h->orbit_ws = nullptr; h->orbit_cap = 0; h->orbit_hdr = 0; h->keycell = nullptr; h->keycell_nkeys = 0; h->keycell_nkeys2 = 0; h->keycell_tag = 0; h->orbit_rec = nullptr; h->orbit_rec_cap = 0; h->field_pipe = nullptr;
    h->g_con = nullptr; h->g4 = nullptr; h->partial = nullptr; h->counter = nullptr;
    const EofGeom& g = h->g;
    cudaError_t e = cudaMalloc(&h->g_con, (size_t)g.nnode * h->gstride * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->g4, (size_t)g.numx * g.numy * 12 * (g.mmax + 1) * 2 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->partial, (size_t)(h->max_ctas + 64) * h->nch_pad * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->counter, 128 * sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemsetAsync(h->counter, 0, 128 * sizeof(unsigned int), stream);
    if (e != cudaSuccess) {                 // a partly built clone must not leak (cudaFree(nullptr) is a no-op)
        bfe_set_cuda_error(e, "bfe_eof_clone");
        bfe_eof_destroy(h);
        return BFE_ERR_CUDA;
    }
    *out = h;
    return BFE_OK;
}

extern "C" void bfe_eof_destroy(bfe_eof* h) {
    if (!h) return;
    if (h->owns_tables) { cudaFree(h->t_acc); if (h->t_force) cudaFree(h->t_force); }
    cudaFree(h->g_con); cudaFree(h->g4); cudaFree(h->partial); cudaFree(h->counter);
    if (h->g4f) cudaFree(h->g4f);
    if (h->orbit_ws) cudaFree(h->orbit_ws);
    if (h->keycell) cudaFree(h->keycell);
    if (h->orbit_rec) cudaFree(h->orbit_rec);
    if (h->sort_ws) cudaFree(h->sort_ws);
    bfe_host_pipe_destroy(h->host_pipe);
    bfe_field_pipe_destroy(h->field_pipe);
    delete h;
}

extern "C" int bfe_eof_accumulate(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                                  const double* mass, double* cos_out, double* sin_out, void* stream_) {
    if (!h || n < 0 || !cos_out || !sin_out) return BFE_ERR_ARG;
    if (n > 0 && (!x || !y || !z || !mass)) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    {
        const bool can_sort = (h->g.mmax <= 6 && h->nch_pad <= 256);
        const int mode = g_bfe_eof_accumulate_mode;
        if (can_sort && (mode == 2 || (mode == 0 && n >= g_bfe_sort_min_particles)))
            return bfe_eof_accumulate_sorted(h, n, x, y, z, mass, cos_out, sin_out, stream);
    }
    const bool fast = (h->g.mmax <= 6 && h->nch_pad <= 256);
    const int tile = fast ? 256 : 64;
    int64_t ntiles = (n + tile - 1) / tile;
    int grid = (int)(ntiles < h->max_ctas ? (ntiles < 1 ? 1 : ntiles) : h->max_ctas);
    if (fast)
        eof_accumulate_kernel<6, 256, 1><<<grid, 256, 0, stream>>>(h->g, h->t_acc, h->nch, h->nch_pad, n, x, y, z,
                                                                  mass, h->partial, h->counter, cos_out, sin_out);
    else
        eof_accumulate_kernel<BFE_MAX_MMAX, 64, 4><<<grid, 256, 0, stream>>>(h->g, h->t_acc, h->nch, h->nch_pad, n, x,
                                                                            y, z, mass, h->partial, h->counter,
                                                                            cos_out, sin_out);
    BFE_LAUNCH_CHECK("eof_accumulate_kernel");
    return BFE_OK;
}

extern "C" int bfe_eof_contract(bfe_eof* h, const double* cosc, const double* sinc, int m1, int m2, int nuse,
                                int no_odd, void* stream_) {
    if (!h || !cosc || !sinc) return BFE_ERR_ARG;
    if (!h->t_force) return BFE_ERR_STATE;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nuse < 0) nuse = 0;
    dim3 grd((h->g.nnode + 255) / 256, (h->g.mmax + 1) * 6);
    const int kt = bfe_kt_begin("eof_contract_kernel", stream);
    BFE_CUDA(bfe_launch(eof_contract_kernel, grd, dim3(128), 0, stream, h->t_force, 6 * h->tab_elems * sizeof(double),
                        h->g, h->t_force, h->tab_elems, cosc, sinc, m1, m2, nuse, no_odd, h->g_con, h->gstride, g_bfe_contract_deep));
    bfe_kt_end(kt, stream);
    BFE_LAUNCH_CHECK("eof_contract_kernel");
    h->contracted = 1;
    h->g4_valid = 0;
    h->g4f_valid = 0;
    return BFE_OK;
}

// per-lane evaluation on the per-cell blocks G4 with 256-bit loads (bfe_eof_eval_blk): half the load instructions
// of eof_force_kernel / eof_points_kernel, no shared-memory stage
template <int MCAP>
__global__ void __launch_bounds__(128)
eof_force_blk_kernel(EofGeom g, const double2* __restrict__ G4, int64_t n,
                     const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                     double* __restrict__ p0, double* __restrict__ p, double* __restrict__ fr,
                     double* __restrict__ fp, double* __restrict__ fz, double* __restrict__ R) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        double r = sqrt(px * px + py * py + 1.e-10);              // eof.py:1070
        EofBin b = bfe_eof_bin(g, r, pz);
        double c1, s1;
        bfe_cossin_phi(px, py, c1, s1);                           // eof.py:1068
        EofField f = bfe_eof_eval_blk<MCAP>(g, G4, b, c1, s1);
        p0[i] = f.p0; p[i] = f.p; fr[i] = f.fr; fp[i] = f.fp; fz[i] = f.fz; R[i] = r;
    }
}

template <int MCAP>
__global__ void __launch_bounds__(128)
eof_points_blk_kernel(EofGeom g, const double2* __restrict__ G4, int64_t n,
                      const double* __restrict__ r, const double* __restrict__ z, const double* __restrict__ phi,
                      double* __restrict__ fr, double* __restrict__ fp, double* __restrict__ fz,
                      double* __restrict__ p, double* __restrict__ p0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        EofBin b = bfe_eof_bin(g, __ldg(r + i), __ldg(z + i));
        double c1, s1;
        sincos(__ldg(phi + i), &s1, &c1);
        EofField f = bfe_eof_eval_blk<MCAP>(g, G4, b, c1, s1);
        fr[i] = f.fr; fp[i] = f.fp; fz[i] = f.fz; p[i] = f.p + f.p0; p0[i] = f.p0;
    }
}

int bfe_eof_ensure_g4f(bfe_eof* h, cudaStream_t stream) {
    if (h->g4f_valid) return BFE_OK;
    if (!h->g4f) BFE_CUDA(cudaMalloc(&h->g4f, (size_t)h->g.numx * h->g.numy * 24 * (h->g.mmax + 1) * sizeof(float)));
    eof_expand_g4f_kernel<<<h->num_sms * 8, 256, 0, stream>>>(h->g, h->g_con, h->gstride, h->g4f);
    BFE_LAUNCH_CHECK("eof_expand_g4f_kernel");
    h->g4f_valid = 1;
    return BFE_OK;
}

int bfe_eof_ensure_g4(bfe_eof* h, cudaStream_t stream) {
    if (h->g4_valid) return BFE_OK;
    eof_expand_g4_kernel<<<h->num_sms * 8, 256, 0, stream>>>(h->g, h->g_con, h->gstride,
                                                            reinterpret_cast<double2*>(h->g4));
    BFE_LAUNCH_CHECK("eof_expand_g4_kernel");
    h->g4_valid = 1;
    return BFE_OK;
}

extern "C" int bfe_eof_force_contracted(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                                        double* p0, double* p, double* fr, double* fp, double* fz, double* R,
                                        void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!x || !y || !z || !p0 || !p || !fr || !fp || !fz || !R) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    {
        const int mode = g_bfe_eof_force_mode;
        if (h->g.mmax <= 6 && (mode == 2 || (mode == 0 && n >= g_bfe_sort_min_particles)))
            return bfe_eof_force_sorted(h, n, x, y, z, p0, p, fr, fp, fz, R, stream);
    }
    int grid = grid_for(n, 128, h->num_sms, 16);
    // EOF alone: the warp-staged copy wins (176 vs 186 us per 10^6 points, profiles/blk_ab.py); the 256-bit per-lane
    // variant is the choice when staging is switched off
    if (h->g.mmax <= 6 && g_bfe_blk_eval && !g_bfe_staged_eval) {
        int rc = bfe_eof_ensure_g4(h, stream);
        if (rc != BFE_OK) return rc;
        eof_force_blk_kernel<6><<<grid, 128, 0, stream>>>(h->g, reinterpret_cast<const double2*>(h->g4), n, x, y, z,
                                                         p0, p, fr, fp, fz, R);
    } else if (h->g.mmax <= 6 && g_bfe_staged_eval) {
        int rc = bfe_eof_ensure_g4(h, stream);
        if (rc != BFE_OK) return rc;
        eof_force_staged_kernel<6><<<grid, 128, 0, stream>>>(h->g, reinterpret_cast<const double2*>(h->g4), n, x, y, z,
                                                            p0, p, fr, fp, fz, R);
    } else if (h->g.mmax <= 6)
        eof_force_kernel<6><<<grid, 128, 0, stream>>>(h->g, h->g_con, h->gstride, n, x, y, z, p0, p, fr, fp, fz, R);
    else
        eof_force_kernel<BFE_MAX_MMAX><<<grid, 128, 0, stream>>>(h->g, h->g_con, h->gstride, n, x, y, z, p0, p, fr,
                                                                 fp, fz, R);
    BFE_LAUNCH_CHECK("eof_force_kernel");
    return BFE_OK;
}

extern "C" int bfe_eof_force(bfe_eof* h, int64_t n, const double* x, const double* y, const double* z,
                             const double* cosc, const double* sinc, int m1, int m2, int nuse, int no_odd,
                             double* p0, double* p, double* fr, double* fp, double* fz, double* R, void* stream) {
    int rc = bfe_eof_contract(h, cosc, sinc, m1, m2, nuse, no_odd, stream);
    if (rc != BFE_OK) return rc;
    return bfe_eof_force_contracted(h, n, x, y, z, p0, p, fr, fp, fz, R, stream);
}

extern "C" int bfe_eof_force_eval_points(bfe_eof* h, int64_t n, const double* r, const double* z, const double* phi,
                                         double* fr, double* fp, double* fz, double* p, double* p0, void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!r || !z || !phi || !fr || !fp || !fz || !p || !p0) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    int grid = grid_for(n, 128, h->num_sms, 16);
    if (h->g.mmax <= 6 && g_bfe_blk_eval && !g_bfe_staged_eval) {
        int rc = bfe_eof_ensure_g4(h, stream);
        if (rc != BFE_OK) return rc;
        eof_points_blk_kernel<6><<<grid, 128, 0, stream>>>(h->g, reinterpret_cast<const double2*>(h->g4), n, r, z, phi,
                                                          fr, fp, fz, p, p0);
    } else if (h->g.mmax <= 6 && g_bfe_staged_eval) {
        int rc = bfe_eof_ensure_g4(h, stream);
        if (rc != BFE_OK) return rc;
        eof_points_staged_kernel<6><<<grid, 128, 0, stream>>>(h->g, reinterpret_cast<const double2*>(h->g4), n, r, z, phi,
                                                             fr, fp, fz, p, p0);
    } else if (h->g.mmax <= 6)
        eof_points_kernel<6><<<grid, 128, 0, stream>>>(h->g, h->g_con, h->gstride, n, r, z, phi, fr, fp, fz, p, p0);
    else
        eof_points_kernel<BFE_MAX_MMAX><<<grid, 128, 0, stream>>>(h->g, h->g_con, h->gstride, n, r, z, phi, fr, fp,
                                                                  fz, p, p0);
    BFE_LAUNCH_CHECK("eof_points_kernel");
    return BFE_OK;
}
