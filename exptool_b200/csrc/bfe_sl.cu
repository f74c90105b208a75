// bfe_sl.cu -- spherical Sturm-Liouville (halo basis) kernels.
//
//   sl_relayout_kernel   : eftable[l][n][i] -> node-major e_node[i][l][n] / sqrt(ev[l][n])
//   sl_accumulate_kernel : spheresl.compute_coefficients_solitary (spheresl.py:567-656)
//   sl_contract_kernel   : A[i][k] = sum_n expcoef[k,n] e_node[i][l(k)][n]
//   sl_force_kernel      : spheresl.all_eval_particles (spheresl.py:1240-1362), potential part
//   sl_points_kernel     : spheresl.force_eval / all_eval (spheresl.py:1107-1234 / 987-1102)
#include "bfe_device.cuh"
#include <math.h>

__global__ void sl_relayout_kernel(SlGeom g, const double* __restrict__ ev, const double* __restrict__ ef,
                                   double* __restrict__ e_node) {
    __shared__ double tile[32][33];
    const int i0 = blockIdx.x * 32, c0 = blockIdx.y * 32;          // c = l*nmax+n
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int c = c0 + r, i = i0 + threadIdx.x;
        double v = 0.0;
        if (c < g.ln && i < g.numr) v = ef[(size_t)c * g.numr + i] / sqrt(ev[c]);   // spheresl.py:332
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int i = i0 + r, c = c0 + threadIdx.x;
        if (i < g.numr && c < g.ln) e_node[(size_t)i * g.ln + c] = tile[threadIdx.x][r];
    }
}

// ---------------------------------------------------------------------------
// accumulate, direct formulation.
//
// Thread (l,n) owns the 2l+1 coefficient rows of its radial function in registers.
// Phase A (thread per particle): r, cos(theta), P_l^m recurrence, cos/sin(m phi)
// recurrence, radial bin; writes w_k = -4pi m f_lm P_lm {1|cos|sin} P0 for all
// (lmax+1)^2 rows k and (i, x1, x2) to shared memory.  Phase B (thread per (l,n)):
// e = x1 E[i][l][n] + x2 E[i+1][l][n] (coalesced over n within l), acc[k] += w_k e.
// ---------------------------------------------------------------------------
template <int LCAP, int TILE>
__global__ void __launch_bounds__(256)
sl_accumulate_kernel(SlGeom g, const double* __restrict__ e_node, const double* __restrict__ xi,
                     const double* __restrict__ p0tab, const double* __restrict__ fac,
                     int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                     const double* __restrict__ z, const double* __restrict__ mass, int no_odd,
                     double* __restrict__ partial, unsigned int* __restrict__ counter,
                     double* __restrict__ expcoef) {
    constexpr int NROW = (LCAP + 1) * (LCAP + 1);
    extern __shared__ double s_dyn[];
    double* s_w = s_dyn;                          // [nrow][TILE]
    double* s_x1 = s_w + (size_t)g.nrow * TILE;   // [TILE]
    double* s_x2 = s_x1 + TILE;                   // [TILE]
    int* s_i = reinterpret_cast<int*>(s_x2 + TILE);
    (void)NROW;
    (void)counter; (void)expcoef;

    const int tid = threadIdx.x;
    const int my_l = tid / g.nmax;                // thread -> (l, n)
    const bool active = tid < g.ln && !(no_odd && (my_l & 1));
    double acc[2 * LCAP + 1];
#pragma unroll
    for (int k = 0; k < 2 * LCAP + 1; ++k) acc[k] = 0.0;

    const int64_t ntiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (tid < TILE) {
            int64_t ip = tile * TILE + tid;
            double px = 0.0, py = 0.0, pz = 0.0, pm = 0.0;
            if (ip < n) { px = __ldg(x + ip); py = __ldg(y + ip); pz = __ldg(z + ip); pm = __ldg(mass + ip); }
            double r2 = BFE_ADD(BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py)), BFE_MUL(pz, pz));   // spheresl.py:610
            double r = fmax(sqrt(r2), 1.0e-10);                   // 611
            double costh = BFE_DIV(pz, r);                        // 612
            double c1, s1;
            bfe_cossin_phi(px, py, c1, s1);                       // 613
            LegTable<LCAP> P;
            bfe_legendre_fast<LCAP>(g.lmax, costh, P);            // 618
            SlBin b = bfe_sl_bin(g, xi, r);                       // 623 -> 309-328
            double P0 = b.x1 * __ldg(p0tab + b.i) + b.x2 * __ldg(p0tab + b.i + 1);
            double W = BFE_FOURPI_NEG * pm * P0;
            s_i[tid] = b.i; s_x1[tid] = b.x1; s_x2[tid] = b.x2;
            // cos/sin(m phi) for m = 0..lmax
            double cm[LCAP + 1], sm[LCAP + 1];
            cm[0] = 1.0; sm[0] = 0.0;
#pragma unroll
            for (int m = 1; m <= LCAP; ++m) {
                cm[m] = cm[m - 1] * c1 - sm[m - 1] * s1;
                sm[m] = sm[m - 1] * c1 + cm[m - 1] * s1;
            }
#pragma unroll
            for (int l = 0; l <= LCAP; ++l) {
                if (l <= g.lmax) {
                    const int k0 = l * l;
                    s_w[(size_t)k0 * TILE + tid] = W * __ldg(fac + l * (g.lmax + 1)) * P.p[l][0];
#pragma unroll
                    for (int m = 1; m <= l; ++m) {
                        double fw = W * __ldg(fac + l * (g.lmax + 1) + m) * P.p[l][m];
                        s_w[(size_t)(k0 + 2 * m - 1) * TILE + tid] = fw * cm[m];
                        s_w[(size_t)(k0 + 2 * m) * TILE + tid] = fw * sm[m];
                    }
                }
            }
        }
        __syncthreads();
        if (active) {
            const double* ecol = e_node + tid;       // tid == l*nmax + n
            const double* wrow = s_w + (size_t)(my_l * my_l) * TILE;
#pragma unroll 2
            for (int p = 0; p < TILE; ++p) {
                const double* e = ecol + (size_t)s_i[p] * g.ln;
                double ev = s_x1[p] * __ldg(e) + s_x2[p] * __ldg(e + g.ln);
#pragma unroll
                for (int k = 0; k < 2 * LCAP + 1; ++k)
                    if (k < 2 * my_l + 1) acc[k] = fma(wrow[(size_t)k * TILE + p], ev, acc[k]);
            }
        }
        __syncthreads();
    }

    // partial layout: [cta][k][n] with k over nrow rows
    const int ncoef = g.nrow * g.nmax;
    if (tid < g.ln) {
        const int nn = tid - my_l * g.nmax;
#pragma unroll
        for (int k = 0; k < 2 * LCAP + 1; ++k)
            if (k < 2 * my_l + 1)
                partial[(size_t)blockIdx.x * ncoef + (size_t)(my_l * my_l + k) * g.nmax + nn] = acc[k];
    }
}

// column sums of the per-CTA partial rows: out[c] = sum_b partial[b][c].  256 threads = 8 columns x 32 row
// slices per CTA; fixed summation order (deterministic).
__global__ void __launch_bounds__(256)
sl_reduce_partials_kernel(const double* __restrict__ partial, int nrows, int ncol, double* __restrict__ out) {
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
}

// A[i][q] (double2, q = l(l+1)/2 + m): .x = cosine row k = l^2 (m=0) or l^2+2m-1, .y = sine row k = l^2+2m.
// A[i][q].{x,y} = sum_{n<nuse} c[k][n] * e_node[i][l][n]; rows outside the l window are zero.
// grid: (ceil(numr/128), nrow); the m=0 sine slots stay zero from create().
__global__ void sl_contract_kernel(SlGeom g, const double* __restrict__ e_node, const double* __restrict__ expcoef,
                                   int l1, int l2, int nuse, int no_odd, double* __restrict__ A, int qstride,
                                   const double* __restrict__ scale_ln) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (i >= g.numr || k >= g.nrow) return;
    int l = (int)sqrt((double)k);
    while ((l + 1) * (l + 1) <= k) ++l;
    while (l * l > k) --l;
    const int r = k - l * l;                       // 0: m=0 ; odd: cos m=(r+1)/2 ; even: sin m=r/2
    const int m = (r + 1) / 2;
    const int comp = (r > 0 && (r & 1) == 0) ? 1 : 0;
    double s = 0.0;
    bool use = (l == 0) || ((l >= l1) && (l <= l2) && !(no_odd && (l & 1)));
    if (use) {
        const double* e = e_node + (size_t)i * g.ln + l * g.nmax;
        const double* c = expcoef + (size_t)k * g.nmax;
        const int nn = nuse < g.nmax ? nuse : g.nmax;
        if (scale_ln) {                            // density rows: e_node * ev = ef * sqrt(ev)   (spheresl.py:148)
            const double* w = scale_ln + l * g.nmax;
            for (int q = 0; q < nn; ++q) s = fma(__ldg(c + q) * __ldg(w + q), __ldg(e + q), s);
        } else {
            for (int q = 0; q < nn; ++q) s = fma(__ldg(c + q), __ldg(e + q), s);
        }
    }
    A[((size_t)i * qstride + (l * (l + 1)) / 2 + m) * 2 + comp] = s;
}

// a_con [i][q = l(l+1)/2 + m] -> A3 [j][3*(offm(m) + l - m) + node], node 0,1,2 = rows j-1, j, j+1 (valid for 1 <= j <= numr-2);
// the block of one j is BFE_A3_STRIDE(npair) double2 long (padded to a multiple of 32 bytes)
__global__ void sl_expand_a3_kernel(SlGeom g, const double2* __restrict__ A, int qstride, double2* __restrict__ A3) {
    const int npair = qstride;
    const int stride = BFE_A3_STRIDE(npair);
    const int64_t total = (int64_t)g.numr * stride;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t / stride), c = (int)(t - (int64_t)j * stride);
        double2 v = make_double2(0.0, 0.0);
        if (c < 3 * npair) {
            const int qm = c / 3, node = c - qm * 3;
            // (m, l) from the column-major pair index qm
            int m = 0, off = 0;
            while (qm >= off + (g.lmax - m + 1)) { off += g.lmax - m + 1; ++m; }
            const int l = m + (qm - off);
            const int row = j - 1 + node;
            if (row >= 0 && row < g.numr) v = A[(size_t)row * qstride + (l * (l + 1)) / 2 + m];
        }
        A3[t] = v;
    }
}

// float blocks of the table_fp32 mode: A3f[i][8*(offm(m) + l - m) + ...], layout and rationale at bfe_sl_eval_blk32
__global__ void sl_expand_a3f_kernel(SlGeom g, const double2* __restrict__ A, int qstride, const double* __restrict__ p0,
                                     float* __restrict__ A3f) {
    const int npair = qstride;
    const int64_t total = (int64_t)(g.numr - 1) * npair;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / npair), qm = (int)(t - (int64_t)i * npair);
        int m = 0, off = 0;
        while (qm >= off + (g.lmax - m + 1)) { off += g.lmax - m + 1; ++m; }
        const int l = m + (qm - off);
        const int q = (l * (l + 1)) / 2 + m;
        const int j = (i == 0) ? 1 : i;
        const double2 alo = A[(size_t)i * qstride + q], ahi = A[(size_t)(i + 1) * qstride + q];
        const double2 am = A[(size_t)(j - 1) * qstride + q], a0 = A[(size_t)j * qstride + q], ap = A[(size_t)(j + 1) * qstride + q];
        const double pm = p0[j - 1], pc = p0[j], pp = p0[j + 1];
        const double umx = pm * am.x, u0x = pc * a0.x, upx = pp * ap.x;
        const double umy = pm * am.y, u0y = pc * a0.y, upy = pp * ap.y;
        float* o = A3f + (size_t)t * 8;
        o[0] = (float)alo.x; o[1] = (float)alo.y; o[2] = (float)ahi.x; o[3] = (float)ahi.y;
        o[4] = (float)(0.5 * (upx - umx)); o[5] = (float)(0.5 * (upy - umy));
        o[6] = (float)((umx - u0x) + (upx - u0x)); o[7] = (float)((umy - u0y) + (upy - u0y));
    }
}

// FP64 polynomial blocks A4 of the per-lane evaluation (layout and rationale at bfe_sl_eval_poly): thread per
// (interval i, entry (m, l) in evaluation order)
__global__ void sl_expand_a4_kernel(SlGeom g, const double2* __restrict__ A, int qstride, const double* __restrict__ p0,
                                    const double* __restrict__ fac, double* __restrict__ A4) {
    const int npair = qstride;
    const int np0 = g.lmax + 1;
    const size_t blk_doubles = (size_t)(4 * np0 + 8 * (npair - np0));
    const int64_t total = (int64_t)(g.numr - 1) * npair;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / npair), qm = (int)(t - (int64_t)i * npair);
        int m = 0, off = 0;
        while (qm >= off + (g.lmax - m + 1)) { off += g.lmax - m + 1; ++m; }
        const int l = m + (qm - off);
        const int q = (l * (l + 1)) / 2 + m;
        const int j = (i == 0) ? 1 : i;
        const double fl = fac[l * (g.lmax + 1) + m];
        const double2 alo = A[(size_t)i * qstride + q], ahi = A[(size_t)(i + 1) * qstride + q];
        const double2 am = A[(size_t)(j - 1) * qstride + q], a0 = A[(size_t)j * qstride + q], ap = A[(size_t)(j + 1) * qstride + q];
        const double pm = p0[j - 1], pc = p0[j], pp = p0[j + 1];
        const double umx = pm * am.x, u0x = pc * a0.x, upx = pp * ap.x;
        const double umy = pm * am.y, u0y = pc * a0.y, upy = pp * ap.y;
        double* o = A4 + (size_t)i * blk_doubles;
        if (m == 0) {
            o += 4 * l;
            o[0] = fl * alo.x; o[1] = fl * (ahi.x - alo.x);
            o[2] = fl * (0.5 * (upx - umx)); o[3] = fl * ((umx - u0x) + (upx - u0x));
        } else {
            o += 4 * np0 + 8 * (qm - np0);
            o[0] = fl * alo.x; o[1] = fl * alo.y; o[2] = fl * (ahi.x - alo.x); o[3] = fl * (ahi.y - alo.y);
            o[4] = fl * (0.5 * (upx - umx)); o[5] = fl * (0.5 * (upy - umy));
            o[6] = fl * ((umx - u0x) + (upx - u0x)); o[7] = fl * ((umy - u0y) + (upy - u0y));
        }
    }
}

// per-lane block evaluation with 256-bit loads (bfe_sl_eval_blk); valid for g.lmax == LCAP
template <int LCAP, bool F32>
__global__ void __launch_bounds__(128)
sl_force_blk_kernel(SlGeom g, const void* __restrict__ A3, const double* __restrict__ xi,
                    const double* __restrict__ p0tab, const double* __restrict__ fac, int64_t n,
                    const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                    double* __restrict__ pot0, double* __restrict__ pot1, double* __restrict__ potr,
                    double* __restrict__ pott, double* __restrict__ potp, double* __restrict__ rr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        double rxy2 = BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py));
        double r = sqrt(BFE_ADD(rxy2, BFE_MUL(pz, pz)));          // spheresl.py:1257 (no epsilon)
        double rxy = sqrt(rxy2);
        double costh = BFE_DIV(pz, r);
        double c1, s1;
        bfe_cossin_phi(px, py, c1, s1);
        SlBin b = bfe_sl_bin(g, xi, r);
        SlField f;
        if constexpr (F32) f = bfe_sl_eval_blk32<LCAP>(g, static_cast<const float*>(A3), p0tab, fac, b, costh, c1, s1, false);
        else               f = bfe_sl_eval_blk<LCAP>(g, A3, p0tab, fac, b, costh, c1, s1, false);
        pot0[i] = f.pot0; pot1[i] = f.pot1; potr[i] = f.potr; pott[i] = f.pott; potp[i] = f.potp; rr[i] = rxy;
    }
}

template <int LCAP>
__global__ void __launch_bounds__(128)
sl_points_blk_kernel(SlGeom g, const void* __restrict__ A3, const double* __restrict__ xi,
                     const double* __restrict__ p0tab, const double* __restrict__ fac, int64_t n,
                     const double* __restrict__ r, const double* __restrict__ costh, const double* __restrict__ phi,
                     int trig_index_l,
                     double* __restrict__ potr, double* __restrict__ pott, double* __restrict__ potp,
                     double* __restrict__ pot1, double* __restrict__ pot0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double c1, s1;
        sincos(__ldg(phi + i), &s1, &c1);
        SlBin b = bfe_sl_bin(g, xi, __ldg(r + i));
        SlField f = bfe_sl_eval_blk<LCAP>(g, A3, p0tab, fac, b, __ldg(costh + i), c1, s1, trig_index_l != 0);
        potr[i] = f.potr; pott[i] = f.pott; potp[i] = f.potp; pot1[i] = f.pot1; pot0[i] = f.pot0;
    }
}

template <int LCAP>
__global__ void __launch_bounds__(128)
sl_force_staged_kernel(SlGeom g, const double2* __restrict__ A3, const double* __restrict__ xi,
                       const double* __restrict__ p0tab, const double* __restrict__ fac, int64_t n,
                       const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                       double* __restrict__ pot0, double* __restrict__ pot1, double* __restrict__ potr,
                       double* __restrict__ pott, double* __restrict__ potp, double* __restrict__ rr) {
    __shared__ double2 s_stage[4][BFE_STAGE_DOUBLE2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* st = s_stage[warp];
    const int64_t wglobal = (int64_t)blockIdx.x * 4 + warp, wtotal = (int64_t)gridDim.x * 4;
    for (int64_t base = wglobal * 32; base < n; base += wtotal * 32) {
        const int64_t i = base + lane;
        const bool on = i < n;
        const int64_t ii = on ? i : n - 1;
        double px = __ldg(x + ii), py = __ldg(y + ii), pz = __ldg(z + ii);
        double rxy2 = BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py));
        double r = sqrt(BFE_ADD(rxy2, BFE_MUL(pz, pz)));          // spheresl.py:1257 (no epsilon)
        double rxy = sqrt(rxy2);
        double costh = BFE_DIV(pz, r);
        double c1, s1;
        bfe_cossin_phi(px, py, c1, s1);
        SlBin b = bfe_sl_bin(g, xi, r);
        SlField f = bfe_sl_eval_staged<LCAP>(g, A3, p0tab, fac, b, costh, c1, s1, false, st, lane);
        if (on) { pot0[i] = f.pot0; pot1[i] = f.pot1; potr[i] = f.potr; pott[i] = f.pott; potp[i] = f.potp; rr[i] = rxy; }
    }
}

template <int LCAP>
__global__ void __launch_bounds__(128)
sl_points_staged_kernel(SlGeom g, const double2* __restrict__ A3, const double* __restrict__ xi,
                        const double* __restrict__ p0tab, const double* __restrict__ fac, int64_t n,
                        const double* __restrict__ r, const double* __restrict__ costh, const double* __restrict__ phi,
                        int trig_index_l,
                        double* __restrict__ potr, double* __restrict__ pott, double* __restrict__ potp,
                        double* __restrict__ pot1, double* __restrict__ pot0) {
    __shared__ double2 s_stage[4][BFE_STAGE_DOUBLE2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* st = s_stage[warp];
    const int64_t wglobal = (int64_t)blockIdx.x * 4 + warp, wtotal = (int64_t)gridDim.x * 4;
    for (int64_t base = wglobal * 32; base < n; base += wtotal * 32) {
        const int64_t i = base + lane;
        const bool on = i < n;
        const int64_t ii = on ? i : n - 1;
        double c1, s1;
        sincos(__ldg(phi + ii), &s1, &c1);
        SlBin b = bfe_sl_bin(g, xi, __ldg(r + ii));
        SlField f = bfe_sl_eval_staged<LCAP>(g, A3, p0tab, fac, b, __ldg(costh + ii), c1, s1, trig_index_l != 0, st, lane);
        if (on) { potr[i] = f.potr; pott[i] = f.pott; potp[i] = f.potp; pot1[i] = f.pot1; pot0[i] = f.pot0; }
    }
}

template <int LCAP>
__global__ void __launch_bounds__(128)
sl_force_kernel(SlGeom g, const double2* __restrict__ A, int kpad, const double* __restrict__ xi,
                const double* __restrict__ p0tab, const double* __restrict__ fac, int64_t n,
                const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                double* __restrict__ pot0, double* __restrict__ pot1, double* __restrict__ potr,
                double* __restrict__ pott, double* __restrict__ potp, double* __restrict__ rr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double px = __ldg(x + i), py = __ldg(y + i), pz = __ldg(z + i);
        double rxy2 = BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py));
        double r = sqrt(BFE_ADD(rxy2, BFE_MUL(pz, pz)));          // spheresl.py:1257 (no epsilon)
        double rxy = sqrt(rxy2);                                  // 1259
        double costh = BFE_DIV(pz, r);                            // 1265
        double c1, s1;
        bfe_cossin_phi(px, py, c1, s1);                           // 1261
        SlBin b = bfe_sl_bin(g, xi, r);
        SlField f = bfe_sl_eval<LCAP>(g, A, kpad, p0tab, fac, b, costh, c1, s1, false);
        pot0[i] = f.pot0; pot1[i] = f.pot1; potr[i] = f.potr; pott[i] = f.pott; potp[i] = f.potp; rr[i] = rxy;
    }
}

template <int LCAP>
__global__ void __launch_bounds__(128)
sl_points_kernel(SlGeom g, const double2* __restrict__ A, int kpad, const double* __restrict__ xi,
                 const double* __restrict__ p0tab, const double* __restrict__ fac, int64_t n,
                 const double* __restrict__ r, const double* __restrict__ costh, const double* __restrict__ phi,
                 int trig_index_l,
                 double* __restrict__ potr, double* __restrict__ pott, double* __restrict__ potp,
                 double* __restrict__ pot1, double* __restrict__ pot0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double c1, s1;
        sincos(__ldg(phi + i), &s1, &c1);
        SlBin b = bfe_sl_bin(g, xi, __ldg(r + i));
        SlField f = bfe_sl_eval<LCAP>(g, A, kpad, p0tab, fac, b, __ldg(costh + i), c1, s1, trig_index_l != 0);
        potr[i] = f.potr; pott[i] = f.pott; potp[i] = f.potp; pot1[i] = f.pot1; pot0[i] = f.pot0;
    }
}

// ---------------------------------------------------------------------------
// Density outputs den0, den1 of spheresl.all_eval (points, 987-1102) and spheresl.all_eval_particles
// (particles, 1240-1362) from contracted density rows  Ad[i][q] = sum_n c[k][n] ef[l,n,i] sqrt(ev[l,n]):
//   dens contribution = (x1 Ad[i] + x2 Ad[i+1]) * (x1 d0[i] + x2 d0[i+1])        (spheresl.py:148)
// The two functions differ and each is reproduced as written:
//   all_eval            den1 starts from the monopole (1046), m=0 term legs[l][0] (1071), densfac 0.25/pi (1092)
//   all_eval_particles  den1 starts from 0 (1271),   m=0 term legs[1][0] (1323),      densfac 0.25*pi (1351)
// POINTS = true: inputs are (r, costh, phi); false: (x, y, z) with r = sqrt(x^2+y^2+z^2) (1257).
// ---------------------------------------------------------------------------
template <int LCAP, bool POINTS>
__global__ void __launch_bounds__(128)
sl_density_kernel(SlGeom g, const double2* __restrict__ Ad, int qstride, const double* __restrict__ xi,
                  const double* __restrict__ d0tab, const double* __restrict__ fac, int64_t n,
                  const double* __restrict__ a, const double* __restrict__ b_, const double* __restrict__ c,
                  double* __restrict__ den0, double* __restrict__ den1) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double r, costh, c1, s1;
        if (POINTS) {
            r = __ldg(a + i); costh = __ldg(b_ + i);
            sincos(__ldg(c + i), &s1, &c1);
        } else {
            const double px = __ldg(a + i), py = __ldg(b_ + i), pz = __ldg(c + i);
            r = sqrt(BFE_ADD(BFE_ADD(BFE_MUL(px, px), BFE_MUL(py, py)), BFE_MUL(pz, pz)));
            costh = BFE_DIV(pz, r);
            bfe_cossin_phi(px, py, c1, s1);
        }
        const SlBin b = bfe_sl_bin(g, xi, r);
        const double D0 = b.x1 * __ldg(d0tab + b.i) + b.x2 * __ldg(d0tab + b.i + 1);
        const double w1 = b.x1 * D0, w2 = b.x2 * D0;
        const double2* row0 = Ad + (size_t)b.i * qstride;
        const double2* row1 = row0 + qstride;
        LegTable<LCAP> T;
        bfe_legendre<LCAP>(g.lmax, costh, T);
        const double mono = __ldg(fac) * (w1 * __ldg(row0).x + w2 * __ldg(row1).x);
        double sum = POINTS ? mono : 0.0;
        const double p10 = (g.lmax >= 1) ? T.p[1][0] : 0.0;
        double cm = 1.0, sm = 0.0;
#pragma unroll
        for (int m = 0; m <= LCAP; ++m) {
            if (m <= g.lmax) {
                if (m > 0) { const double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1; cm = cn; sm = sn; }
#pragma unroll
                for (int l = (m > 1 ? m : 1); l <= LCAP; ++l) {
                    if (l <= g.lmax) {
                        const int q = (l * (l + 1)) / 2 + m;
                        const double2 a0 = __ldg(row0 + q), a1 = __ldg(row1 + q);
                        const double fl = __ldg(fac + l * (g.lmax + 1) + m);
                        const double sc = w1 * a0.x + w2 * a1.x;
                        if (m == 0) {
                            sum += fl * (POINTS ? T.p[l][0] : p10) * sc;
                        } else {
                            const double ss = w1 * a0.y + w2 * a1.y;
                            sum += fl * T.p[l][m] * (sc * cm + ss * sm);
                        }
                    }
                }
            }
        }
        const double densfac = POINTS ? 0.25 / M_PI : 0.25 * M_PI;
        den0[i] = mono * densfac;
        den1[i] = sum * densfac;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int sl_grid_for(int64_t n, int block, int num_sms, int per_sm) {
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)num_sms * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

template <int LCAP, int TILE>
static int sl_acc_launch(bfe_sl* h, int64_t n, const double* x, const double* y, const double* z,
                         const double* mass, int no_odd, double* expcoef, cudaStream_t stream) {
    int block = (h->g.ln + 31) / 32 * 32;
    if (block < TILE) block = TILE;
    if (block > 256) return BFE_ERR_UNSUPPORTED;
    size_t smem = ((size_t)h->g.nrow * TILE + 2 * TILE) * sizeof(double) + TILE * sizeof(int);
    auto kern = sl_accumulate_kernel<LCAP, TILE>;
    static size_t attr_smem = 0;                  // one per template instance; grows with nrow
    if (smem > attr_smem) {
        BFE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    int64_t ntiles = (n + TILE - 1) / TILE;
    int grid = (int)(ntiles < h->max_ctas ? (ntiles < 1 ? 1 : ntiles) : h->max_ctas);
    kern<<<grid, block, smem, stream>>>(h->g, h->e_node, h->xi, h->p0, h->fac, n, x, y, z, mass, no_odd,
                                        h->partial, h->counter, expcoef);
    BFE_LAUNCH_CHECK("sl_accumulate_kernel");
    const int ncoef = h->g.nrow * h->g.nmax;
    sl_reduce_partials_kernel<<<(ncoef + 7) / 8, 256, 0, stream>>>(h->partial, grid, ncoef, expcoef);
    BFE_LAUNCH_CHECK("sl_reduce_partials_kernel");
    return BFE_OK;
}

extern "C" int bfe_sl_create(const bfe_sl_params* p, const double* evtable, const double* eftable,
                             const double* xi, const double* p0, const double* d0, void* stream_, bfe_sl** out) {
    if (!p || !out || !evtable || !eftable || !xi || !p0) return BFE_ERR_ARG;
    if (p->lmax < 0 || p->lmax > BFE_MAX_LMAX || p->nmax < 1 || p->numr < 3) return BFE_ERR_ARG;
    if (p->cmap != 0 && p->cmap != 1) return BFE_ERR_UNSUPPORTED;   // cmap=2 is broken in the reference
    cudaStream_t stream = (cudaStream_t)stream_;
    bfe_sl* h = new bfe_sl();
    h->par = *p;
    SlGeom& g = h->g;
    g.lmax = p->lmax; g.nmax = p->nmax; g.numr = p->numr; g.cmap = p->cmap; g.scale = p->scale;
    g.nrow = (p->lmax + 1) * (p->lmax + 1);
    g.ln = (p->lmax + 1) * p->nmax;
    if (g.ln > 256) { delete h; return BFE_ERR_UNSUPPORTED; }
    BFE_CUDA(cudaGetDevice(&h->device));
    BFE_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device));
    double xi01[2];
    BFE_CUDA(cudaMemcpyAsync(xi01, xi, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    BFE_CUDA(cudaStreamSynchronize(stream));
    g.xi0 = xi01[0];                        // np.min(xi), spheresl.py:319
    g.dxi = xi01[1] - xi01[0];              // xi[1]-xi[0], spheresl.py:317
    g.inv_scale = 1.0 / g.scale; g.inv_dxi = 1.0 / g.dxi;
    h->kpad = (p->lmax + 1) * (p->lmax + 2) / 2;     // (l,m) pairs per radial node, one double2 each
    h->contracted = 0;
    h->max_ctas = h->num_sms * 6;
    size_t nr = (size_t)p->numr;
    BFE_CUDA(cudaMalloc(&h->e_node, nr * g.ln * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->xi, nr * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->p0, nr * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->d0, nr * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->ev, (size_t)g.ln * sizeof(double)));
    h->have_d0 = d0 ? 1 : 0;
    h->ad_con = nullptr;
    h->dens_contracted = 0;
    BFE_CUDA(cudaMalloc(&h->fac, sizeof(double) * (p->lmax + 1) * (p->lmax + 1)));
    BFE_CUDA(cudaMalloc(&h->a_con, nr * h->kpad * 2 * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->a3, nr * (size_t)BFE_A3_STRIDE(h->kpad) * 2 * sizeof(double)));
    h->a3_valid = 0;
    h->a3f = nullptr; h->a3f_valid = 0;
    h->a4 = nullptr; h->a4_valid = 0;
    BFE_CUDA(cudaMalloc(&h->partial, (size_t)h->max_ctas * g.nrow * g.nmax * sizeof(double)));
    BFE_CUDA(cudaMalloc(&h->counter, 4 * sizeof(unsigned int)));
    BFE_CUDA(cudaMemsetAsync(h->counter, 0, 4 * sizeof(unsigned int), stream));
    h->sort_cap = 0; h->sort_ws = nullptr; h->table_fp32 = -1; h->host_pipe = nullptr;
    BFE_CUDA(cudaMemsetAsync(h->a_con, 0, nr * h->kpad * 2 * sizeof(double), stream));
    BFE_CUDA(cudaMemcpyAsync(h->xi, xi, nr * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    BFE_CUDA(cudaMemcpyAsync(h->p0, p0, nr * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    if (d0) BFE_CUDA(cudaMemcpyAsync(h->d0, d0, nr * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    BFE_CUDA(cudaMemcpyAsync(h->ev, evtable, (size_t)g.ln * sizeof(double), cudaMemcpyDeviceToDevice, stream));
    // factorial_return, spheresl.py:823-863 (lgamma for scipy.special.gammaln)
    for (int l = 0; l <= p->lmax; ++l)
        for (int m = 0; m <= p->lmax; ++m) {
            double v = 0.0;
            if (m <= l) {
                v = sqrt((0.5 * l + 0.25) / M_PI * exp(lgamma(1.0 + l - m) - lgamma(1.0 + l + m)));
                if (m != 0) v *= sqrt(2.0);
            }
            h->fac_host[l * (p->lmax + 1) + m] = v;
        }
    BFE_CUDA(cudaMemcpyAsync(h->fac, h->fac_host, sizeof(double) * (p->lmax + 1) * (p->lmax + 1),
                             cudaMemcpyHostToDevice, stream));
    dim3 blk(32, 8), grd((p->numr + 31) / 32, (g.ln + 31) / 32);
    sl_relayout_kernel<<<grd, blk, 0, stream>>>(g, evtable, eftable, h->e_node);
    BFE_LAUNCH_CHECK("sl_relayout_kernel");
    BFE_CUDA(cudaStreamSynchronize(stream));      // fac_host is pageable; keep create() self-contained
    *out = h;
    return BFE_OK;
}

extern "C" void bfe_sl_destroy(bfe_sl* h) {
    if (!h) return;
    cudaFree(h->e_node); cudaFree(h->xi); cudaFree(h->p0); cudaFree(h->d0); cudaFree(h->fac);
    cudaFree(h->ev); if (h->ad_con) cudaFree(h->ad_con); if (h->a3f) cudaFree(h->a3f); if (h->a4) cudaFree(h->a4);
    cudaFree(h->a_con); cudaFree(h->a3); cudaFree(h->partial); cudaFree(h->counter);
    if (h->sort_ws) cudaFree(h->sort_ws);
    bfe_host_pipe_destroy(h->host_pipe);
    delete h;
}

extern "C" int bfe_sl_accumulate(bfe_sl* h, int64_t n, const double* x, const double* y, const double* z,
                                 const double* mass, int no_odd, double* expcoef, void* stream_) {
    if (!h || n < 0 || !expcoef) return BFE_ERR_ARG;
    if (n > 0 && (!x || !y || !z || !mass)) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    {
        const int mode = g_bfe_sl_accumulate_mode;
        // the sorted pipeline has ~60 us of fixed cost (4 launches): it wins from a few 1e5 particles up
        if (bfe_sl_sorted_supported(h) && (mode == 2 || (mode == 0 && n >= 8 * (int64_t)g_bfe_sort_min_particles)))
            return bfe_sl_accumulate_sorted(h, n, x, y, z, mass, no_odd, expcoef, stream);
    }
    if (h->g.lmax <= 4) return sl_acc_launch<4, 128>(h, n, x, y, z, mass, no_odd, expcoef, stream);
    if (h->g.lmax <= 6) return sl_acc_launch<6, 128>(h, n, x, y, z, mass, no_odd, expcoef, stream);
    return sl_acc_launch<BFE_MAX_LMAX, 32>(h, n, x, y, z, mass, no_odd, expcoef, stream);
}

extern "C" int bfe_sl_contract(bfe_sl* h, const double* expcoef, int l1, int l2, int nuse, int no_odd,
                               void* stream_) {
    if (!h || !expcoef) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nuse < 0) nuse = 0;
    dim3 grd((h->g.numr + 127) / 128, h->g.nrow);
    sl_contract_kernel<<<grd, 128, 0, stream>>>(h->g, h->e_node, expcoef, l1, l2, nuse, no_odd, h->a_con, h->kpad, nullptr);
    BFE_LAUNCH_CHECK("sl_contract_kernel");
    h->contracted = 1;
    h->a3_valid = 0;
    h->a3f_valid = 0;
    h->a4_valid = 0;
    return BFE_OK;
}

int bfe_sl_ensure_a3f(bfe_sl* h, cudaStream_t stream) {
    if (h->a3f_valid) return BFE_OK;
    if (!h->a3f) BFE_CUDA(cudaMalloc(&h->a3f, (size_t)h->g.numr * BFE_A3F_STRIDE(h->kpad) * sizeof(float)));
    sl_expand_a3f_kernel<<<h->num_sms * 4, 256, 0, stream>>>(h->g, reinterpret_cast<const double2*>(h->a_con), h->kpad, h->p0, h->a3f);
    BFE_LAUNCH_CHECK("sl_expand_a3f_kernel");
    h->a3f_valid = 1;
    return BFE_OK;
}

int bfe_sl_ensure_a4(bfe_sl* h, cudaStream_t stream) {
    if (h->a4_valid) return BFE_OK;
    if (!h->a4) BFE_CUDA(cudaMalloc(&h->a4, (size_t)h->g.numr * BFE_A4_BYTES(h->g.lmax)));
    sl_expand_a4_kernel<<<h->num_sms * 4, 256, 0, stream>>>(h->g, reinterpret_cast<const double2*>(h->a_con), h->kpad, h->p0,
                                                           h->fac, h->a4);
    BFE_LAUNCH_CHECK("sl_expand_a4_kernel");
    h->a4_valid = 1;
    return BFE_OK;
}

int bfe_sl_ensure_a3(bfe_sl* h, cudaStream_t stream) {
    if (h->a3_valid) return BFE_OK;
    sl_expand_a3_kernel<<<h->num_sms * 4, 256, 0, stream>>>(h->g, reinterpret_cast<const double2*>(h->a_con), h->kpad,
                                                           reinterpret_cast<double2*>(h->a3));
    BFE_LAUNCH_CHECK("sl_expand_a3_kernel");
    h->a3_valid = 1;
    return BFE_OK;
}

#define SL_DISPATCH(KERN, ...)                                                                   \
    do {                                                                                         \
        if (h->g.lmax <= 4)      KERN<4><<<grid, 128, 0, stream>>>(__VA_ARGS__);                  \
        else if (h->g.lmax <= 6) KERN<6><<<grid, 128, 0, stream>>>(__VA_ARGS__);                  \
        else                     KERN<BFE_MAX_LMAX><<<grid, 128, 0, stream>>>(__VA_ARGS__);       \
    } while (0)

extern "C" int bfe_sl_force_contracted(bfe_sl* h, int64_t n, const double* x, const double* y, const double* z,
                                       double* pot0, double* pot1, double* potr, double* pott, double* potp,
                                       double* rr, void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!x || !y || !z || !pot0 || !pot1 || !potr || !pott || !potp || !rr) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    int grid = sl_grid_for(n, 128, h->num_sms, 16);
    if (g_bfe_blk_eval && (h->g.lmax == 4 || h->g.lmax == 6)) {
        const bool f32 = bfe_use_fp32(h);
        int rc = f32 ? bfe_sl_ensure_a3f(h, stream) : bfe_sl_ensure_a4(h, stream);
        if (rc != BFE_OK) return rc;
        const void* A3 = f32 ? (const void*)h->a3f : (const void*)h->a4;      // FP64: the polynomial blocks
#define SL_BLK(L, F) sl_force_blk_kernel<L, F><<<grid, 128, 0, stream>>>(h->g, A3, h->xi, h->p0, h->fac, n, x, y, z, pot0, pot1, potr, pott, potp, rr)
        if (h->g.lmax == 4) { if (f32) SL_BLK(4, true); else SL_BLK(4, false); }
        else                { if (f32) SL_BLK(6, true); else SL_BLK(6, false); }
#undef SL_BLK
        BFE_LAUNCH_CHECK("sl_force_blk_kernel");
        return BFE_OK;
    }
    if (g_bfe_staged_eval && (h->g.lmax == 4 || h->g.lmax == 6)) {
        int rc = bfe_sl_ensure_a3(h, stream);
        if (rc != BFE_OK) return rc;
        const double2* A3 = reinterpret_cast<const double2*>(h->a3);
        if (h->g.lmax == 4)
            sl_force_staged_kernel<4><<<grid, 128, 0, stream>>>(h->g, A3, h->xi, h->p0, h->fac, n, x, y, z, pot0, pot1,
                                                               potr, pott, potp, rr);
        else
            sl_force_staged_kernel<6><<<grid, 128, 0, stream>>>(h->g, A3, h->xi, h->p0, h->fac, n, x, y, z, pot0, pot1,
                                                               potr, pott, potp, rr);
        BFE_LAUNCH_CHECK("sl_force_staged_kernel");
        return BFE_OK;
    }
    SL_DISPATCH(sl_force_kernel, h->g, reinterpret_cast<const double2*>(h->a_con), h->kpad, h->xi, h->p0, h->fac, n, x, y, z, pot0, pot1, potr, pott,
                potp, rr);
    BFE_LAUNCH_CHECK("sl_force_kernel");
    return BFE_OK;
}

extern "C" int bfe_sl_force(bfe_sl* h, int64_t n, const double* x, const double* y, const double* z,
                            const double* expcoef, int l1, int l2, int no_odd,
                            double* pot0, double* pot1, double* potr, double* pott, double* potp, double* rr,
                            void* stream) {
    if (!h) return BFE_ERR_ARG;
    int rc = bfe_sl_contract(h, expcoef, l1, l2, h->g.nmax, no_odd, stream);
    if (rc != BFE_OK) return rc;
    return bfe_sl_force_contracted(h, n, x, y, z, pot0, pot1, potr, pott, potp, rr, stream);
}

extern "C" int bfe_sl_force_eval_points(bfe_sl* h, int64_t n, const double* r, const double* costh,
                                        const double* phi, int trig_index_l,
                                        double* potr, double* pott, double* potp, double* pot1, double* pot0,
                                        void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!r || !costh || !phi || !potr || !pott || !potp || !pot1 || !pot0) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    int grid = sl_grid_for(n, 128, h->num_sms, 16);
    if (g_bfe_blk_eval && (h->g.lmax == 4 || h->g.lmax == 6)) {
        int rc = bfe_sl_ensure_a4(h, stream);
        if (rc != BFE_OK) return rc;
        const void* A3 = (const void*)h->a4;
        if (h->g.lmax == 4) sl_points_blk_kernel<4><<<grid, 128, 0, stream>>>(h->g, A3, h->xi, h->p0, h->fac, n, r, costh, phi, trig_index_l, potr, pott, potp, pot1, pot0);
        else                sl_points_blk_kernel<6><<<grid, 128, 0, stream>>>(h->g, A3, h->xi, h->p0, h->fac, n, r, costh, phi, trig_index_l, potr, pott, potp, pot1, pot0);
        BFE_LAUNCH_CHECK("sl_points_blk_kernel");
        return BFE_OK;
    }
    if (g_bfe_staged_eval && (h->g.lmax == 4 || h->g.lmax == 6)) {
        int rc = bfe_sl_ensure_a3(h, stream);
        if (rc != BFE_OK) return rc;
        const double2* A3 = reinterpret_cast<const double2*>(h->a3);
        if (h->g.lmax == 4)
            sl_points_staged_kernel<4><<<grid, 128, 0, stream>>>(h->g, A3, h->xi, h->p0, h->fac, n, r, costh, phi,
                                                                trig_index_l, potr, pott, potp, pot1, pot0);
        else
            sl_points_staged_kernel<6><<<grid, 128, 0, stream>>>(h->g, A3, h->xi, h->p0, h->fac, n, r, costh, phi,
                                                                trig_index_l, potr, pott, potp, pot1, pot0);
        BFE_LAUNCH_CHECK("sl_points_staged_kernel");
        return BFE_OK;
    }
    SL_DISPATCH(sl_points_kernel, h->g, reinterpret_cast<const double2*>(h->a_con), h->kpad, h->xi, h->p0, h->fac, n, r, costh, phi, trig_index_l,
                potr, pott, potp, pot1, pot0);
    BFE_LAUNCH_CHECK("sl_points_kernel");
    return BFE_OK;
}

// ---------------------------------------------------------------------------
// density outputs
// ---------------------------------------------------------------------------
extern "C" int bfe_sl_contract_density(bfe_sl* h, const double* expcoef, int l1, int l2, int nuse, int no_odd,
                                       void* stream_) {
    if (!h || !expcoef) return BFE_ERR_ARG;
    if (!h->have_d0) return BFE_ERR_STATE;           // created without the model density table
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!h->ad_con) {
        const size_t bytes = (size_t)h->g.numr * h->kpad * 2 * sizeof(double);
        BFE_CUDA(cudaMalloc(&h->ad_con, bytes));
        BFE_CUDA(cudaMemsetAsync(h->ad_con, 0, bytes, stream));     // m=0 sine slots stay zero
    }
    if (nuse < 0) nuse = 0;
    dim3 grd((h->g.numr + 127) / 128, h->g.nrow);
    sl_contract_kernel<<<grd, 128, 0, stream>>>(h->g, h->e_node, expcoef, l1, l2, nuse, no_odd, h->ad_con, h->kpad, h->ev);
    BFE_LAUNCH_CHECK("sl_contract_kernel");
    h->dens_contracted = 1;
    return BFE_OK;
}

template <bool POINTS>
static int sl_density_launch(bfe_sl* h, int64_t n, const double* a, const double* b, const double* c,
                             double* den0, double* den1, cudaStream_t stream) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (!h->dens_contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!a || !b || !c || !den0 || !den1) return BFE_ERR_ARG;
    int grid = sl_grid_for(n, 128, h->num_sms, 16);
    const double2* Ad = reinterpret_cast<const double2*>(h->ad_con);
    if (h->g.lmax <= 4)      sl_density_kernel<4, POINTS><<<grid, 128, 0, stream>>>(h->g, Ad, h->kpad, h->xi, h->d0, h->fac, n, a, b, c, den0, den1);
    else if (h->g.lmax <= 6) sl_density_kernel<6, POINTS><<<grid, 128, 0, stream>>>(h->g, Ad, h->kpad, h->xi, h->d0, h->fac, n, a, b, c, den0, den1);
    else                     sl_density_kernel<BFE_MAX_LMAX, POINTS><<<grid, 128, 0, stream>>>(h->g, Ad, h->kpad, h->xi, h->d0, h->fac, n, a, b, c, den0, den1);
    BFE_LAUNCH_CHECK("sl_density_kernel");
    return BFE_OK;
}

extern "C" int bfe_sl_density_contracted(bfe_sl* h, int64_t n, const double* x, const double* y, const double* z,
                                         double* den0, double* den1, void* stream) {
    return sl_density_launch<false>(h, n, x, y, z, den0, den1, (cudaStream_t)stream);
}

extern "C" int bfe_sl_density_eval_points(bfe_sl* h, int64_t n, const double* r, const double* costh,
                                          const double* phi, double* den0, double* den1, void* stream) {
    return sl_density_launch<true>(h, n, r, costh, phi, den0, den1, (cudaStream_t)stream);
}
