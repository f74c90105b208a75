// bfe_field.cu -- combined disc+halo force and leapfrog orbit integration.
//
//   field_cart_kernel : Fields.return_forces_cart (potential.py:445-497), thread per point
//   leapfrog_kernel   : integrate.leapfrog_integrate (integrate.py:53-190), thread per orbit,
//                       phase-space state in registers for the whole integration
#include "bfe_device.cuh"
#include <string.h>
#include <cstdio>
#include <atomic>

template <int MCAP, int LCAP, bool CYL>
__global__ void __launch_bounds__(128)
field_cart_kernel(EofGeom ge, const double* __restrict__ G, int gstride,
                  SlGeom gs, const double2* __restrict__ A, int kpad, const double* __restrict__ xi,
                  const double* __restrict__ p0tab, const double* __restrict__ fac,
                  int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                  const double* __restrict__ z, double crot, double srot, double* __restrict__ out8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        CartForce f = bfe_field_cart<MCAP, LCAP, CYL>(ge, G, gstride, gs, A, kpad, xi, p0tab, fac,
                                                      __ldg(x + i), __ldg(y + i), __ldg(z + i), crot, srot);
        out8[i] = f.fxd;
        out8[n + i] = f.fxh;
        out8[2 * n + i] = f.fyd;
        out8[3 * n + i] = f.fyh;
        out8[4 * n + i] = f.fzd;
        out8[5 * n + i] = f.fzh;
        out8[6 * n + i] = f.pd;
        out8[7 * n + i] = f.ph;
    }
}

// warp-cooperative variants (rows through the shared-memory stage; valid for mmax <= MCAP, lmax == LCAP)
template <int MCAP, int LCAP, bool CYL>
__global__ void __launch_bounds__(128)
field_cart_staged_kernel(EofGeom ge, const double2* __restrict__ G4, SlGeom gs, const double2* __restrict__ A3,
                         const double* __restrict__ xi, const double* __restrict__ p0tab, const double* __restrict__ fac,
                         int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                         const double* __restrict__ z, double crot, double srot, double* __restrict__ out8) {
    __shared__ double2 s_stage[4][BFE_STAGE_DOUBLE2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* st = s_stage[warp];
    const int64_t wglobal = (int64_t)blockIdx.x * 4 + warp, wtotal = (int64_t)gridDim.x * 4;
    for (int64_t base = wglobal * 32; base < n; base += wtotal * 32) {
        const int64_t i = base + lane;
        const bool on = i < n;
        const int64_t ii = on ? i : n - 1;
        CartForce f = bfe_field_cart_staged<MCAP, LCAP, CYL>(ge, G4, gs, A3, xi, p0tab, fac, __ldg(x + ii), __ldg(y + ii),
                                                             __ldg(z + ii), crot, srot, st, lane);
        if (on) {
            out8[i] = f.fxd; out8[n + i] = f.fxh; out8[2 * n + i] = f.fyd; out8[3 * n + i] = f.fyh;
            out8[4 * n + i] = f.fzd; out8[5 * n + i] = f.fzh; out8[6 * n + i] = f.pd; out8[7 * n + i] = f.ph;
        }
    }
}

template <int MCAP, int LCAP>
__global__ void __launch_bounds__(128)
leapfrog_kernel(EofGeom ge, const double* __restrict__ G, int gstride,
                SlGeom gs, const double2* __restrict__ A, int kpad, const double* __restrict__ xi,
                const double* __restrict__ p0tab, const double* __restrict__ fac,
                int64_t norbit, int64_t nint, double dt, const double* __restrict__ dt_orbit, double rotfreq,
                double* __restrict__ state6, double* __restrict__ traj, int64_t traj_stride,
                int apse, int ap_max, int* __restrict__ nsteps_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= norbit) return;
    if (dt_orbit) dt = dt_orbit[i];                          // per-orbit step (integrate_grid, integrate.py:871)
    double px = state6[i], py = state6[norbit + i], pz = state6[2 * norbit + i];
    double vx = state6[3 * norbit + i], vy = state6[4 * norbit + i], vz = state6[5 * norbit + i];
    const double w = BFE_TWOPI * rotfreq;                    // barpos = 2 pi rotfreq (k dt), integrate.py:94-97
    const double hdt2 = 0.5 * (dt * dt);

    double srot, crot;
    sincos(w * (0.0 * dt), &srot, &crot);
    CartForce f = bfe_field_cart<MCAP, LCAP>(ge, G, gstride, gs, A, kpad, xi, p0tab, fac, px, py, pz, crot, srot);
    double ax = f.fxd + f.fxh, ay = f.fyd + f.fyh, az = f.fzd + f.fzh, pot = f.pd + f.ph;   // 119-122
    if (traj) {
        double* t = traj + i;
        t[0] = px; t[norbit] = py; t[2 * norbit] = pz; t[3 * norbit] = vx; t[4 * norbit] = vy; t[5 * norbit] = vz;
        t[6 * norbit] = pot; t[7 * norbit] = ax; t[8 * norbit] = ay; t[9 * norbit] = az;
    }
    int n_aps = 0;
    double rs0 = 0.0, rs1 = px * px + py * py;               // planar r^2 at steps k-2, k-1
    int64_t step = 1;
    while (n_aps < ap_max && step < nint) {                  // 126
        px = px + (vx * dt) + (ax * hdt2);                   // 129-131
        py = py + (vy * dt) + (ay * hdt2);
        pz = pz + (vz * dt) + (az * hdt2);
        sincos(w * ((double)step * dt), &srot, &crot);
        f = bfe_field_cart<MCAP, LCAP>(ge, G, gstride, gs, A, kpad, xi, p0tab, fac, px, py, pz, crot, srot);
        double bx = f.fxd + f.fxh, by = f.fyd + f.fyh, bz = f.fzd + f.fzh;                   // 134-138
        pot = f.pd + f.ph;
        vx = vx + (0.5 * (ax + bx) * dt);                    // 141-143
        vy = vy + (0.5 * (ay + by) * dt);
        vz = vz + (0.5 * (az + bz) * dt);
        ax = bx; ay = by; az = bz;
        double rs2 = px * px + py * py;
        if (apse && step > 1 && rs1 > rs0 && rs1 > rs2) ++n_aps;                              // 146-153
        rs0 = rs1; rs1 = rs2;
        if (traj && (step % traj_stride) == 0) {
            double* t = traj + (step / traj_stride) * 10 * norbit + i;
            t[0] = px; t[norbit] = py; t[2 * norbit] = pz; t[3 * norbit] = vx; t[4 * norbit] = vy;
            t[5 * norbit] = vz; t[6 * norbit] = pot; t[7 * norbit] = ax; t[8 * norbit] = ay; t[9 * norbit] = az;
        }
        ++step;
    }
    state6[i] = px; state6[norbit + i] = py; state6[2 * norbit + i] = pz;
    state6[3 * norbit + i] = vx; state6[4 * norbit + i] = vy; state6[5 * norbit + i] = vz;
    if (nsteps_out) nsteps_out[i] = (int)step;
}

// per-lane variants on the per-cell / per-interval blocks with 256-bit loads (bfe_field_cart_blk): half the load
// instructions of the kernels above, no shared-memory stage; valid for mmax <= MCAP, lmax == LCAP == 6
template <int MCAP, int LCAP, bool CYL, bool F32>
__global__ void __launch_bounds__(128)
field_cart_blk_kernel(EofGeom ge, const void* __restrict__ G4, SlGeom gs, const void* __restrict__ A3,
                      const double* __restrict__ xi, const double* __restrict__ p0tab, const SlFacP fac,
                      int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                      const double* __restrict__ z, double crot, double srot, double* __restrict__ out8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        CartForce f = bfe_field_cart_blk<MCAP, LCAP, CYL, F32>(ge, G4, gs, A3, xi, p0tab, fac,
                                                               __ldg(x + i), __ldg(y + i), __ldg(z + i), crot, srot);
        out8[i] = f.fxd; out8[n + i] = f.fxh; out8[2 * n + i] = f.fyd; out8[3 * n + i] = f.fyh;
        out8[4 * n + i] = f.fzd; out8[5 * n + i] = f.fzh; out8[6 * n + i] = f.pd; out8[7 * n + i] = f.ph;
    }
}

template <int MCAP, int LCAP, bool F32>
__global__ void __launch_bounds__(128)
leapfrog_blk_kernel(EofGeom ge, const void* __restrict__ G4, SlGeom gs, const void* __restrict__ A3,
                    const double* __restrict__ xi, const double* __restrict__ p0tab, const SlFacP fac,
                    int64_t norbit, int64_t nint, double dt, const double* __restrict__ dt_orbit, double rotfreq,
                    double* __restrict__ state6, double* __restrict__ traj, int64_t traj_stride,
                    int apse, int ap_max, int* __restrict__ nsteps_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= norbit) return;
    if (dt_orbit) dt = dt_orbit[i];
    double px = state6[i], py = state6[norbit + i], pz = state6[2 * norbit + i];
    double vx = state6[3 * norbit + i], vy = state6[4 * norbit + i], vz = state6[5 * norbit + i];
    const double w = BFE_TWOPI * rotfreq;
    const double hdt2 = 0.5 * (dt * dt);
    double srot, crot;
    sincos(w * (0.0 * dt), &srot, &crot);
    CartForce f = bfe_field_cart_blk<MCAP, LCAP, false, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
    double ax = f.fxd + f.fxh, ay = f.fyd + f.fyh, az = f.fzd + f.fzh, pot = f.pd + f.ph;
    if (traj) {
        double* t = traj + i;
        t[0] = px; t[norbit] = py; t[2 * norbit] = pz; t[3 * norbit] = vx; t[4 * norbit] = vy; t[5 * norbit] = vz;
        t[6 * norbit] = pot; t[7 * norbit] = ax; t[8 * norbit] = ay; t[9 * norbit] = az;
    }
    int n_aps = 0;
    double rs0 = 0.0, rs1 = px * px + py * py;
    int64_t step = 1;
    while (n_aps < ap_max && step < nint) {
        px = px + (vx * dt) + (ax * hdt2);
        py = py + (vy * dt) + (ay * hdt2);
        pz = pz + (vz * dt) + (az * hdt2);
        sincos(w * ((double)step * dt), &srot, &crot);
        f = bfe_field_cart_blk<MCAP, LCAP, false, F32>(ge, G4, gs, A3, xi, p0tab, fac, px, py, pz, crot, srot);
        double bx = f.fxd + f.fxh, by = f.fyd + f.fyh, bz = f.fzd + f.fzh;
        pot = f.pd + f.ph;
        vx = vx + (0.5 * (ax + bx) * dt);
        vy = vy + (0.5 * (ay + by) * dt);
        vz = vz + (0.5 * (az + bz) * dt);
        ax = bx; ay = by; az = bz;
        double rs2 = px * px + py * py;
        if (apse && step > 1 && rs1 > rs0 && rs1 > rs2) ++n_aps;
        rs0 = rs1; rs1 = rs2;
        if (traj && (step % traj_stride) == 0) {
            double* t = traj + (step / traj_stride) * 10 * norbit + i;
            t[0] = px; t[norbit] = py; t[2 * norbit] = pz; t[3 * norbit] = vx; t[4 * norbit] = vy;
            t[5 * norbit] = vz; t[6 * norbit] = pot; t[7 * norbit] = ax; t[8 * norbit] = ay; t[9 * norbit] = az;
        }
        ++step;
    }
    state6[i] = px; state6[norbit + i] = py; state6[2 * norbit + i] = pz;
    state6[3 * norbit + i] = vx; state6[4 * norbit + i] = vy; state6[5 * norbit + i] = vz;
    if (nsteps_out) nsteps_out[i] = (int)step;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#define FIELD_DISPATCH(KERN, ...)                                                                         \
    do {                                                                                                  \
        if (he->g.mmax <= 6 && hs->g.lmax <= 4)      KERN<6, 4><<<grid, 128, 0, stream>>>(__VA_ARGS__);    \
        else if (he->g.mmax <= 6 && hs->g.lmax <= 6) KERN<6, 6><<<grid, 128, 0, stream>>>(__VA_ARGS__);    \
        else KERN<BFE_MAX_MMAX, BFE_MAX_LMAX><<<grid, 128, 0, stream>>>(__VA_ARGS__);                      \
    } while (0)

#define FIELD_DISPATCH_CYL(CYL, ...)                                                                               \
    do {                                                                                                          \
        if (he->g.mmax <= 6 && hs->g.lmax <= 4)      field_cart_kernel<6, 4, CYL><<<grid, 128, 0, stream>>>(__VA_ARGS__); \
        else if (he->g.mmax <= 6 && hs->g.lmax <= 6) field_cart_kernel<6, 6, CYL><<<grid, 128, 0, stream>>>(__VA_ARGS__); \
        else field_cart_kernel<BFE_MAX_MMAX, BFE_MAX_LMAX, CYL><<<grid, 128, 0, stream>>>(__VA_ARGS__);            \
    } while (0)

static int field_force_impl(bfe_eof* he, bfe_sl* hs, int64_t n, const double* x, const double* y, const double* z,
                            double rotpos, double* out8, void* stream_, bool cyl) {
    if (!he || !hs || n < 0) return BFE_ERR_ARG;
    if (!he->contracted || !hs->contracted) return BFE_ERR_STATE;
    if (n == 0) return BFE_OK;
    if (!x || !y || !z || !out8) return BFE_ERR_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    int64_t need = (n + 127) / 128, cap = (int64_t)he->num_sms * 16;
    int grid = (int)(need < cap ? need : cap);
    double crot = cos(rotpos), srot = sin(rotpos);
    // large point sets: evaluated in table-cell order (bfe_orbit_sort.cu), results returned in the caller's order
    if (g_bfe_blk_eval && he->g.mmax <= 6 && (hs->g.lmax == 4 || hs->g.lmax == 6) && g_bfe_field_sort_min > 0 &&
        n >= g_bfe_field_sort_min && n < ((int64_t)1 << 31))
        return bfe_field_force_sorted(he, hs, n, x, y, z, crot, srot, out8, cyl, stream);
    if (g_bfe_blk_eval && he->g.mmax <= 6 && (hs->g.lmax == 4 || hs->g.lmax == 6)) {
        const bool f32 = bfe_use_fp32(he);
        int rc = f32 ? bfe_eof_ensure_g4f(he, stream) : bfe_eof_ensure_g4(he, stream);
        if (rc == BFE_OK) rc = f32 ? bfe_sl_ensure_a3f(hs, stream) : bfe_sl_ensure_a4(hs, stream);
        if (rc != BFE_OK) return rc;
        const void* G4 = f32 ? (const void*)he->g4f : (const void*)he->g4;
        const void* A3 = f32 ? (const void*)hs->a3f : (const void*)hs->a4;      // FP64: the polynomial blocks
        const SlFacP facp = bfe_sl_facp(hs);
#define FIELD_BLK(L, C, F) field_cart_blk_kernel<6, L, C, F><<<grid, 128, 0, stream>>>(he->g, G4, hs->g, A3, hs->xi, hs->p0, \
                                                                                    facp, n, x, y, z, crot, srot, out8)
#define FIELD_BLK2(L, C) do { if (f32) FIELD_BLK(L, C, true); else FIELD_BLK(L, C, false); } while (0)
        if (hs->g.lmax == 4) { if (cyl) FIELD_BLK2(4, true); else FIELD_BLK2(4, false); }
        else                 { if (cyl) FIELD_BLK2(6, true); else FIELD_BLK2(6, false); }
#undef FIELD_BLK2
#undef FIELD_BLK
        BFE_LAUNCH_CHECK("field_cart_blk_kernel");
        return BFE_OK;
    }
    if (g_bfe_staged_eval && he->g.mmax <= 6 && (hs->g.lmax == 4 || hs->g.lmax == 6)) {
        int rc = bfe_eof_ensure_g4(he, stream);
        if (rc == BFE_OK) rc = bfe_sl_ensure_a3(hs, stream);
        if (rc != BFE_OK) return rc;
        const double2* G4 = reinterpret_cast<const double2*>(he->g4);
        const double2* A3 = reinterpret_cast<const double2*>(hs->a3);
#define FIELD_STAGED(L, C) field_cart_staged_kernel<6, L, C><<<grid, 128, 0, stream>>>(he->g, G4, hs->g, A3, hs->xi, hs->p0, \
                                                                                 hs->fac, n, x, y, z, crot, srot, out8)
        if (hs->g.lmax == 4) { if (cyl) FIELD_STAGED(4, true); else FIELD_STAGED(4, false); }
        else                 { if (cyl) FIELD_STAGED(6, true); else FIELD_STAGED(6, false); }
#undef FIELD_STAGED
        BFE_LAUNCH_CHECK("field_cart_staged_kernel");
        return BFE_OK;
    }
    if (cyl)
        FIELD_DISPATCH_CYL(true, he->g, he->g_con, he->gstride, hs->g, reinterpret_cast<const double2*>(hs->a_con), hs->kpad, hs->xi, hs->p0, hs->fac, n,
                           x, y, z, crot, srot, out8);
    else
        FIELD_DISPATCH_CYL(false, he->g, he->g_con, he->gstride, hs->g, reinterpret_cast<const double2*>(hs->a_con), hs->kpad, hs->xi, hs->p0, hs->fac,
                           n, x, y, z, crot, srot, out8);
    BFE_LAUNCH_CHECK("field_cart_kernel");
    return BFE_OK;
}

extern "C" int bfe_field_force_cart(bfe_eof* he, bfe_sl* hs, int64_t n, const double* x, const double* y,
                                    const double* z, double rotpos, double* out8, void* stream) {
    return field_force_impl(he, hs, n, x, y, z, rotpos, out8, stream, false);
}

extern "C" int bfe_field_force_cyl(bfe_eof* he, bfe_sl* hs, int64_t n, const double* x, const double* y,
                                   const double* z, double rotpos, double* out8, void* stream) {
    return field_force_impl(he, hs, n, x, y, z, rotpos, out8, stream, true);
}

static int leapfrog_impl(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, double dt, const double* dt_orbit,
                         double rotfreq, double* state6, double* traj, int64_t traj_stride, int apse, int ap_max,
                         int32_t* nsteps_out, void* stream_) {
    if (!he || !hs || norbit < 0 || nint < 1) return BFE_ERR_ARG;
    if (!he->contracted || !hs->contracted) return BFE_ERR_STATE;
    if (norbit == 0) return BFE_OK;
    if (!state6) return BFE_ERR_ARG;
    if (traj && traj_stride < 1) return BFE_ERR_ARG;
    if (ap_max < 0) ap_max = 0;          // ap_max = 0: the reference's loop (integrate.py:126) takes no step at all
    cudaStream_t stream = (cudaStream_t)stream_;
    int grid = (int)((norbit + 127) / 128);
    // large batches without trajectory / apocentre bookkeeping: orbits kept cell-coherent by a re-sort every K steps
    if (g_bfe_blk_eval && he->g.mmax <= 6 && (hs->g.lmax == 4 || hs->g.lmax == 6) && !traj && !apse && ap_max > 0 &&
        g_bfe_orbit_resort > 0 && norbit >= g_bfe_orbit_sort_min && norbit < ((int64_t)1 << 31) &&
        nint > 2 * (int64_t)g_bfe_orbit_resort + 2 && nint < ((int64_t)1 << 31))
        return bfe_leapfrog_sorted(he, hs, norbit, nint, dt, dt_orbit, rotfreq, state6, nsteps_out, stream);
    if (g_bfe_blk_eval && he->g.mmax <= 6 && (hs->g.lmax == 4 || hs->g.lmax == 6)) {
        const bool f32 = bfe_use_fp32(he);
        int rc = f32 ? bfe_eof_ensure_g4f(he, stream) : bfe_eof_ensure_g4(he, stream);
        if (rc == BFE_OK) rc = f32 ? bfe_sl_ensure_a3f(hs, stream) : bfe_sl_ensure_a4(hs, stream);
        if (rc != BFE_OK) return rc;
        const void* G4 = f32 ? (const void*)he->g4f : (const void*)he->g4;
        const void* A3 = f32 ? (const void*)hs->a3f : (const void*)hs->a4;      // FP64: the polynomial blocks
        const SlFacP facp = bfe_sl_facp(hs);
#define LEAP_BLK(L, F) leapfrog_blk_kernel<6, L, F><<<grid, 128, 0, stream>>>(he->g, G4, hs->g, A3, hs->xi, hs->p0, facp, norbit, \
                                                                           nint, dt, dt_orbit, rotfreq, state6, traj, traj_stride, apse, ap_max, nsteps_out)
        if (hs->g.lmax == 4) { if (f32) LEAP_BLK(4, true); else LEAP_BLK(4, false); }
        else                 { if (f32) LEAP_BLK(6, true); else LEAP_BLK(6, false); }
#undef LEAP_BLK
        BFE_LAUNCH_CHECK("leapfrog_blk_kernel");
        return BFE_OK;
    }
    // the per-lane kernel is used for orbits: consecutive steps of one orbit re-read the same table rows, which the
    // per-lane loads find in L1, whereas the warp-staged variant re-copies them every step (measured 15 % slower)
    FIELD_DISPATCH(leapfrog_kernel, he->g, he->g_con, he->gstride, hs->g, reinterpret_cast<const double2*>(hs->a_con), hs->kpad, hs->xi, hs->p0,
                   hs->fac, norbit, nint, dt, dt_orbit, rotfreq, state6, traj, traj_stride, apse, ap_max, nsteps_out);
    BFE_LAUNCH_CHECK("leapfrog_kernel");
    return BFE_OK;
}

extern "C" int bfe_leapfrog(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, double dt, double rotfreq,
                            double* state6, double* traj, int64_t traj_stride, int apse, int ap_max,
                            int32_t* nsteps_out, void* stream) {
    return leapfrog_impl(he, hs, norbit, nint, dt, nullptr, rotfreq, state6, traj, traj_stride, apse, ap_max,
                         nsteps_out, stream);
}

extern "C" int bfe_leapfrog_dt(bfe_eof* he, bfe_sl* hs, int64_t norbit, int64_t nint, const double* dt_orbit,
                               double rotfreq, double* state6, double* traj, int64_t traj_stride, int apse,
                               int ap_max, int32_t* nsteps_out, void* stream) {
    if (!dt_orbit) return BFE_ERR_ARG;
    return leapfrog_impl(he, hs, norbit, nint, 0.0, dt_orbit, rotfreq, state6, traj, traj_stride, apse, ap_max,
                         nsteps_out, stream);
}

// ---------------------------------------------------------------------------
// library-wide helpers
// ---------------------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};
static thread_local char g_cuda_err[256] = "";

int g_bfe_eof_accumulate_mode = 0;
int g_bfe_eof_force_mode = 0;
int g_bfe_sort_min_particles = 32768;
int g_bfe_sl_accumulate_mode = 0;
int g_bfe_sl_deposit_mode = 0;
int g_bfe_staged_eval = 1;
int g_bfe_blk_eval = 1;
int g_bfe_table_fp32 = 0;
int g_bfe_force_mma = 1;
static int g_bfe_time_kernels = 0;
int g_bfe_pdl = 1;
int g_bfe_grid_pct = 100;
int g_bfe_contract_deep = 0;
int g_bfe_l2_persist = 0;
size_t g_bfe_l2_window_max = 0;

// persisting-L2 set-aside for the table windows (device-wide limit, set once per process and device)
int bfe_l2_persist_setup(size_t want_bytes) {
    static int done_dev = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return BFE_ERR_CUDA;
    if (done_dev == dev) return BFE_OK;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return BFE_ERR_CUDA;
    size_t cap = (size_t)prop.persistingL2CacheMaxSize;
    g_bfe_l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    size_t sz = want_bytes < cap ? want_bytes : cap;
    if (sz > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, sz) != cudaSuccess) { cudaGetLastError(); return BFE_ERR_CUDA; }
    done_dev = dev;
    return BFE_OK;
}

extern "C" int bfe_set_option(const char* name, int value) {
    if (!name) return BFE_ERR_ARG;
    if (!strcmp(name, "eof_accumulate_mode")) { g_bfe_eof_accumulate_mode = value; return BFE_OK; }
    if (!strcmp(name, "eof_force_mode")) { g_bfe_eof_force_mode = value; return BFE_OK; }
    if (!strcmp(name, "time_kernels")) { g_bfe_time_kernels = value; return BFE_OK; }
    if (!strcmp(name, "staged_eval")) { g_bfe_staged_eval = value; return BFE_OK; }
    if (!strcmp(name, "blk_eval")) { g_bfe_blk_eval = value; return BFE_OK; }
    if (!strcmp(name, "table_fp32")) { g_bfe_table_fp32 = value; return BFE_OK; }
    if (!strcmp(name, "force_mma")) { g_bfe_force_mma = value; return BFE_OK; }
    if (!strcmp(name, "pdl")) { g_bfe_pdl = value; return BFE_OK; }
    if (!strcmp(name, "field_sort_chunk")) { g_bfe_field_sort_chunk = value; return BFE_OK; }
    if (!strcmp(name, "field_sort_min")) { g_bfe_field_sort_min = value; return BFE_OK; }
    if (!strcmp(name, "stage_eval")) { g_bfe_stage_eval = value; return BFE_OK; }
    if (!strcmp(name, "sort_stable")) { g_bfe_sort_stable = value; return BFE_OK; }
    if (!strcmp(name, "key_subbits")) { g_bfe_key_subbits = value; return BFE_OK; }
    if (!strcmp(name, "key_mode")) { g_bfe_key_mode = value & 3; return BFE_OK; }
    if (!strcmp(name, "field_eval_static")) { g_bfe_field_eval_static = value ? 1 : 0; return BFE_OK; }
    if (!strcmp(name, "field_gather_stream")) { g_bfe_field_gather_stream = value ? 1 : 0; return BFE_OK; }
    if (!strcmp(name, "field_support_slim")) { g_bfe_field_support_slim = value < 0 ? 0 : (value > 4 ? 4 : value); return BFE_OK; }
    if (!strcmp(name, "orbit_key_subbits")) { g_bfe_orbit_key_subbits = value; return BFE_OK; }
    if (!strcmp(name, "orbit_resort")) { g_bfe_orbit_resort = value; return BFE_OK; }
    if (!strcmp(name, "orbit_sort_min")) { g_bfe_orbit_sort_min = value; return BFE_OK; }
    if (!strcmp(name, "grid_pct")) { if (value < 1 || value > 100) return BFE_ERR_ARG; g_bfe_grid_pct = value; return BFE_OK; }
    if (!strcmp(name, "contract_deep")) { g_bfe_contract_deep = value; return BFE_OK; }
    if (!strcmp(name, "sl_flush_cost")) { g_bfe_sl_flush_cost = value < 0 ? 0 : value; return BFE_OK; }
    if (!strcmp(name, "host_chunk")) { g_bfe_host_chunk = value; return BFE_OK; }
    if (!strcmp(name, "host_reuse")) { g_bfe_host_reuse = value; return BFE_OK; }
    if (!strcmp(name, "host_threads")) { if (value < 0 || value > 64) return BFE_ERR_ARG; g_bfe_host_threads = value; return BFE_OK; }
    if (!strcmp(name, "l2_persist")) {
        g_bfe_l2_persist = value;
        if (value > 0) return bfe_l2_persist_setup((size_t)96 << 20);
        return BFE_OK;
    }
    if (!strcmp(name, "sl_accumulate_mode")) { g_bfe_sl_accumulate_mode = value; return BFE_OK; }
    if (!strcmp(name, "sl_deposit_mode")) { g_bfe_sl_deposit_mode = value; return BFE_OK; }
    if (!strcmp(name, "sort_min_particles")) { g_bfe_sort_min_particles = value; return BFE_OK; }
    return BFE_ERR_ARG;
}

// current value of a runtime option (so that tests and tools can restore what they change); INT_MIN for unknown names
extern "C" int bfe_get_option(const char* name) {
    if (!name) return -2147483647 - 1;
    if (!strcmp(name, "eof_accumulate_mode")) return g_bfe_eof_accumulate_mode;
    if (!strcmp(name, "eof_force_mode")) return g_bfe_eof_force_mode;
    if (!strcmp(name, "time_kernels")) return g_bfe_time_kernels;
    if (!strcmp(name, "staged_eval")) return g_bfe_staged_eval;
    if (!strcmp(name, "blk_eval")) return g_bfe_blk_eval;
    if (!strcmp(name, "table_fp32")) return g_bfe_table_fp32;
    if (!strcmp(name, "force_mma")) return g_bfe_force_mma;
    if (!strcmp(name, "pdl")) return g_bfe_pdl;
    if (!strcmp(name, "field_sort_chunk")) return g_bfe_field_sort_chunk;
    if (!strcmp(name, "field_sort_min")) return g_bfe_field_sort_min;
    if (!strcmp(name, "stage_eval")) return g_bfe_stage_eval;
    if (!strcmp(name, "sort_stable")) return g_bfe_sort_stable;
    if (!strcmp(name, "key_subbits")) return g_bfe_key_subbits;
    if (!strcmp(name, "key_mode")) return g_bfe_key_mode;
    if (!strcmp(name, "field_eval_static")) return g_bfe_field_eval_static;
    if (!strcmp(name, "field_gather_stream")) return g_bfe_field_gather_stream;
    if (!strcmp(name, "field_support_slim")) return g_bfe_field_support_slim;
    if (!strcmp(name, "keycell_nkeys")) return g_bfe_keycell_nkeys_last;
    if (!strcmp(name, "orbit_key_subbits")) return g_bfe_orbit_key_subbits;
    if (!strcmp(name, "orbit_resort")) return g_bfe_orbit_resort;
    if (!strcmp(name, "orbit_sort_min")) return g_bfe_orbit_sort_min;
    if (!strcmp(name, "grid_pct")) return g_bfe_grid_pct;
    if (!strcmp(name, "contract_deep")) return g_bfe_contract_deep;
    if (!strcmp(name, "sl_flush_cost")) return g_bfe_sl_flush_cost;
    if (!strcmp(name, "host_chunk")) return g_bfe_host_chunk;
    if (!strcmp(name, "host_reuse")) return g_bfe_host_reuse;
    if (!strcmp(name, "host_threads")) return g_bfe_host_threads;
    if (!strcmp(name, "host_reused_last")) return g_bfe_host_reused_last;
    if (!strcmp(name, "l2_persist")) return g_bfe_l2_persist;
    if (!strcmp(name, "sl_accumulate_mode")) return g_bfe_sl_accumulate_mode;
    if (!strcmp(name, "sl_deposit_mode")) return g_bfe_sl_deposit_mode;
    if (!strcmp(name, "sort_min_particles")) return g_bfe_sort_min_particles;
    return -2147483647 - 1;
}

// per-handle table precision of the per-point field kernels (-1: follow the process option "table_fp32")
extern "C" int bfe_eof_set_table_fp32(bfe_eof* h, int value) {
    if (!h || value < -1 || value > 1) return BFE_ERR_ARG;
    h->table_fp32 = value;
    return BFE_OK;
}
extern "C" int bfe_sl_set_table_fp32(bfe_sl* h, int value) {
    if (!h || value < -1 || value > 1) return BFE_ERR_ARG;
    h->table_fp32 = value;
    return BFE_OK;
}

// ---- optional per-kernel event timing
struct BfeKernelTimer { char name[48]; cudaEvent_t a, b; bool made, used; };
static BfeKernelTimer g_kt[24];
static int g_kt_n = 0;

int bfe_kt_begin(const char* name, cudaStream_t stream) {
    if (!g_bfe_time_kernels) return -1;
    int slot = -1;
    for (int i = 0; i < g_kt_n; ++i) if (!strcmp(g_kt[i].name, name)) { slot = i; break; }
    if (slot < 0) {
        if (g_kt_n >= 24) return -1;
        slot = g_kt_n++;
        strncpy(g_kt[slot].name, name, 47); g_kt[slot].name[47] = 0;
        g_kt[slot].made = false; g_kt[slot].used = false;
    }
    if (!g_kt[slot].made) {
        if (cudaEventCreate(&g_kt[slot].a) != cudaSuccess || cudaEventCreate(&g_kt[slot].b) != cudaSuccess) return -1;
        g_kt[slot].made = true;
    }
    cudaEventRecord(g_kt[slot].a, stream);
    return slot;
}

void bfe_kt_end(int slot, cudaStream_t stream) {
    if (slot < 0) return;
    cudaEventRecord(g_kt[slot].b, stream);
    g_kt[slot].used = true;
}

// duration (ms) of the most recent launch of `name` recorded while "time_kernels" was on; < 0 if none.
extern "C" double bfe_kernel_time_ms(const char* name) {
    if (!name) return -1.0;
    for (int i = 0; i < g_kt_n; ++i)
        if (!strcmp(g_kt[i].name, name) && g_kt[i].used) {
            float ms = -1.f;
            if (cudaEventSynchronize(g_kt[i].b) != cudaSuccess) return -1.0;
            if (cudaEventElapsedTime(&ms, g_kt[i].a, g_kt[i].b) != cudaSuccess) return -1.0;
            return (double)ms;
        }
    return -1.0;
}

extern "C" void bfe_count_launch(int n) { g_launches.fetch_add((uint64_t)n); }
extern "C" uint64_t bfe_launch_count(void) { return g_launches.load(); }
extern "C" int bfe_version(void) { return 100; }

void bfe_set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", where, cudaGetErrorString(e));
}
extern "C" const char* bfe_last_cuda_error(void) { return g_cuda_err; }

extern "C" const char* bfe_error_string(int code) {
    switch (code) {
        case BFE_OK: return "ok";
        case BFE_ERR_ARG: return "invalid argument";
        case BFE_ERR_CUDA: return "CUDA error (see bfe_last_cuda_error)";
        case BFE_ERR_UNSUPPORTED: return "basis size or mapping not supported by the compiled kernels";
        case BFE_ERR_STATE: return "handle holds no coefficient contraction (call bfe_*_contract first)";
        default: return "unknown error";
    }
}
