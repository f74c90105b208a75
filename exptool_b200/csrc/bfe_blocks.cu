// bfe_blocks.cu -- the per-point building blocks of the path as callable entry points.
//
// The hot kernels evaluate these inline (bfe_device.cuh); the reference also exposes them as functions
// (SURVEY.md section 8a rows a5, a6, a14, a15, a16), so they are exported with the same meaning:
//
//   eof_bins_kernel      : eof.return_bins (eof.py:354-427)                 -> X, Y, ix, iy
//   eof_get_pot_kernel   : eof.get_pot (eof.py:430-457)                      -> Vc, Vs (mmax+1, norder, n)
//   sl_radial_kernel     : spheresl.get_halo_dens_pot_force (spheresl.py:106-160), get_halo_pot_matrix (301-335)
//                                                                            -> dens, force, pot (lmax+1, nmax, n)
//   legendre_kernel      : spheresl.legendre_R / dlegendre_R (spheresl.py:664-770) -> P, dP (lmax+1, lmax+1, n)
//
// Outputs are point-minor (the reference's trailing particle axis), written coalesced along the points.
#include "bfe_device.cuh"

__global__ void __launch_bounds__(256)
eof_bins_kernel(EofGeom g, int64_t n, const double* __restrict__ r, const double* __restrict__ z,
                double* __restrict__ X, double* __restrict__ Y, long long* __restrict__ ix, long long* __restrict__ iy) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double xv = (bfe_r_to_xi(__ldg(r + i), g.cmap, g.inv_ascale) - g.xmin) * g.inv_dx;      // eof.py:394
        double yv = (bfe_z_to_y(__ldg(z + i), g.inv_hscale) - g.ymin) * g.inv_dy;               // 395
        int jx = (int)xv, jy = (int)yv;                                                     // truncation, 404-405
        if (jx < 0) jx = 0;                                                                 // 410
        if (xv < 0.0) xv = 0.0;                                                             // 412
        if (jx >= g.numx) jx = g.numx - 1;                                                  // 414; 415 is a no-op: X stays
        if (jy < 0) jy = 0;
        if (yv < 0.0) yv = 0.0;
        if (jy >= g.numy) jy = g.numy - 1;
        X[i] = xv; Y[i] = yv; ix[i] = jx; iy[i] = jy;
    }
}

// block (32 points, 8 channel lanes): every thread computes the bin of its point once and then walks the
// channels ty, ty+8, ...; stores are coalesced along the points.
__global__ void __launch_bounds__(256)
eof_get_pot_kernel(EofGeom g, const double* __restrict__ t_acc, int nch_pad, int64_t n,
                   const double* __restrict__ r, const double* __restrict__ z, double fac,
                   double* __restrict__ Vc, double* __restrict__ Vs) {
    const int ncos = (g.mmax + 1) * g.norder;
    const int nsin = g.mmax * g.norder;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < n; base += (int64_t)gridDim.x * 32) {
        const int64_t i = base + threadIdx.x;
        if (i >= n) continue;
        const EofBin b = bfe_eof_bin(g, __ldg(r + i), __ldg(z + i));
        const double* t00 = t_acc + (size_t)b.node * nch_pad;
        const double* t01 = t00 + nch_pad;                       // (ix, iy+1)
        const double* t10 = t00 + (size_t)g.ny1 * nch_pad;       // (ix+1, iy)
        const double* t11 = t10 + nch_pad;
        for (int c = threadIdx.y; c < ncos + nsin; c += blockDim.y) {
            // eof.py:453-455: fac * (T00 c00 + T10 c10 + T01 c01 + T11 c11), left to right
            const double v = fac * (__ldg(t00 + c) * b.c00 + __ldg(t10 + c) * b.c10 + __ldg(t01 + c) * b.c01 +
                                    __ldg(t11 + c) * b.c11);
            if (c < ncos) Vc[(size_t)c * n + i] = v;
            else          Vs[(size_t)(c - ncos + g.norder) * n + i] = v;
        }
        for (int c = threadIdx.y; c < g.norder; c += blockDim.y) Vs[(size_t)c * n + i] = 0.0;   // m = 0 sine plane (eof.py:293)
    }
}

__global__ void __launch_bounds__(256)
sl_radial_kernel(SlGeom g, const double* __restrict__ e_node, const double* __restrict__ ev,
                 const double* __restrict__ xi, const double* __restrict__ p0, const double* __restrict__ d0,
                 int64_t n, const double* __restrict__ r,
                 double* __restrict__ dens, double* __restrict__ force, double* __restrict__ pot) {
    for (int64_t base = (int64_t)blockIdx.x * 32; base < n; base += (int64_t)gridDim.x * 32) {
        const int64_t i = base + threadIdx.x;
        if (i >= n) continue;
        const SlBin b = bfe_sl_bin(g, xi, __ldg(r + i));
        const int j = (b.i == 0) ? 1 : b.i;                                  // spheresl.py:150-153
        const double P0 = b.x1 * __ldg(p0 + b.i) + b.x2 * __ldg(p0 + b.i + 1);
        const double D0 = d0 ? b.x1 * __ldg(d0 + b.i) + b.x2 * __ldg(d0 + b.i + 1) : 0.0;
        const double fa = b.fac * (b.x2 - 0.5) * __ldg(p0 + j - 1), fb = b.fac * (-2.0 * b.x2) * __ldg(p0 + j),
                     fc = b.fac * (b.x2 + 0.5) * __ldg(p0 + j + 1);
        const double* e0 = e_node + (size_t)b.i * g.ln;
        const double* e1 = e0 + g.ln;
        const double* ea = e_node + (size_t)(j - 1) * g.ln;
        for (int c = threadIdx.y; c < g.ln; c += blockDim.y) {
            const double lin = b.x1 * __ldg(e0 + c) + b.x2 * __ldg(e1 + c);      // (x1 ef_i + x2 ef_{i+1}) / sqrt(ev)
            if (pot) pot[(size_t)c * n + i] = lin * P0;                          // 157-158 / 333
            if (dens) dens[(size_t)c * n + i] = lin * __ldg(ev + c) * D0;        // 148: ... * sqrt(ev) * d0
            if (force) force[(size_t)c * n + i] = fa * __ldg(ea + c) + fb * __ldg(ea + g.ln + c) + fc * __ldg(ea + 2 * g.ln + c);   // 153-155
        }
    }
}

template <int LCAP>
__global__ void __launch_bounds__(128)
legendre_kernel(int lmax, int64_t n, const double* __restrict__ x, double* __restrict__ P, double* __restrict__ dP) {
    const int l1 = lmax + 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double xv = __ldg(x + i);
        LegTable<LCAP> T, D;
        bfe_legendre<LCAP>(lmax, xv, T);
#pragma unroll
        for (int l = 0; l <= LCAP; ++l)
#pragma unroll
            for (int m = 0; m <= LCAP; ++m)
                if (l <= lmax && m <= lmax) {
                    double v = (m <= l) ? T.p[l][m] : 0.0;
                    if (!isfinite(v)) v = 0.0;                               // spheresl.py:698 / 750
                    if (m <= l) T.p[l][m] = v;
                    P[(size_t)(l * l1 + m) * n + i] = v;
                }
        if (dP) {
            bfe_dlegendre<LCAP>(lmax, xv, T, D);
#pragma unroll
            for (int l = 0; l <= LCAP; ++l)
#pragma unroll
                for (int m = 0; m <= LCAP; ++m)
                    if (l <= lmax && m <= lmax) {
                        double v = (m <= l) ? D.p[l][m] : 0.0;
                        if (!isfinite(v)) v = 0.0;                           // 768
                        dP[(size_t)(l * l1 + m) * n + i] = v;
                    }
        }
    }
}

static int blocks_grid(int64_t n, int per, int cap) {
    int64_t need = (n + per - 1) / per;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

extern "C" int bfe_eof_return_bins(const bfe_eof_params* p, int64_t n, const double* r, const double* z,
                                   double* X, double* Y, long long* ix, long long* iy, void* stream_) {
    if (!p || n < 0) return BFE_ERR_ARG;
    if (n == 0) return BFE_OK;
    if (!r || !z || !X || !Y || !ix || !iy) return BFE_ERR_ARG;
    if (p->dx == 0.0 || p->dy == 0.0) return BFE_ERR_ARG;
    EofGeom g;
    g.mmax = p->mmax; g.norder = p->norder; g.numx = p->numx; g.numy = p->numy; g.cmap = p->cmap;
    g.ny1 = p->numy + 1; g.nnode = (p->numx + 1) * (p->numy + 1);
    g.xmin = p->xmin; g.dx = p->dx; g.ymin = p->ymin; g.dy = p->dy; g.ascale = p->ascale; g.hscale = p->hscale;
    g.inv_dx = 1.0 / p->dx; g.inv_dy = 1.0 / p->dy;
    g.inv_ascale = 1.0 / p->ascale; g.inv_hscale = 1.0 / p->hscale;
    eof_bins_kernel<<<blocks_grid(n, 256, 148 * 8), 256, 0, (cudaStream_t)stream_>>>(g, n, r, z, X, Y, ix, iy);
    BFE_LAUNCH_CHECK("eof_bins_kernel");
    return BFE_OK;
}

extern "C" int bfe_eof_get_pot(bfe_eof* h, int64_t n, const double* r, const double* z, double fac,
                               double* Vc, double* Vs, void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (n == 0) return BFE_OK;
    if (!r || !z || !Vc || !Vs) return BFE_ERR_ARG;
    eof_get_pot_kernel<<<blocks_grid(n, 32, h->num_sms * 8), dim3(32, 8), 0, (cudaStream_t)stream_>>>(
        h->g, h->t_acc, h->nch_pad, n, r, z, fac, Vc, Vs);
    BFE_LAUNCH_CHECK("eof_get_pot_kernel");
    return BFE_OK;
}

extern "C" int bfe_sl_radial_matrices(bfe_sl* h, int64_t n, const double* r, double* dens, double* force,
                                      double* pot, void* stream_) {
    if (!h || n < 0) return BFE_ERR_ARG;
    if (n == 0) return BFE_OK;
    if (!r || (!dens && !force && !pot)) return BFE_ERR_ARG;
    if (dens && !h->have_d0) return BFE_ERR_STATE;
    sl_radial_kernel<<<blocks_grid(n, 32, h->num_sms * 8), dim3(32, 8), 0, (cudaStream_t)stream_>>>(
        h->g, h->e_node, h->ev, h->xi, h->p0, dens ? h->d0 : nullptr, n, r, dens, force, pot);
    BFE_LAUNCH_CHECK("sl_radial_kernel");
    return BFE_OK;
}

extern "C" int bfe_legendre_tables(int lmax, int64_t n, const double* x, double* P, double* dP, void* stream_) {
    if (lmax < 0 || lmax > BFE_MAX_LMAX || n < 0) return BFE_ERR_ARG;
    if (n == 0) return BFE_OK;
    if (!x || !P) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = blocks_grid(n, 128, 148 * 8);
    if (lmax <= 6) legendre_kernel<6><<<grid, 128, 0, stream>>>(lmax, n, x, P, dP);
    else           legendre_kernel<BFE_MAX_LMAX><<<grid, 128, 0, stream>>>(lmax, n, x, P, dP);
    BFE_LAUNCH_CHECK("legendre_kernel");
    return BFE_OK;
}
