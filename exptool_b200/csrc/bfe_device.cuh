// bfe_device.cuh -- per-point device arithmetic shared by all kernels.
//
// Every function restates one piece of the reference (file:line cited); the
// oracle (oracle/oracle_np.py) restates the same lines on the CPU.
#pragma once
#include "bfe_internal.h"

// programmatic dependent launch (see bfe_launch, bfe_internal.h): no-ops when the kernel was launched without
// the attribute.  bfe_pdl_wait() must precede the first global-memory access of a kernel of the chain.
__device__ __forceinline__ void bfe_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bfe_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#define BFE_FOURPI_NEG (-12.566370614359172953850573533118)
#define BFE_TWOPI      (6.283185307179586476925286766559)

// ---------------------------------------------------------------------------
// column sum over the per-CTA partial rows, executed by the last CTA to finish.
// Rows are summed in a fixed order (deterministic); 16 independent L2 loads are kept in flight
// per thread -- a naive one-load-at-a-time loop over ~600 rows costs ~50 us of pure L2 latency.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double bfe_column_sum(const double* __restrict__ partial, int nrows, int stride, int col) {
    const double* q = partial + col;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int b = 0;
    for (; b + 15 < nrows; b += 16) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcg(q + (size_t)(b + u) * stride);
#pragma unroll
        for (int u = 0; u < 16; u += 4) { s0 += v[u]; s1 += v[u + 1]; s2 += v[u + 2]; s3 += v[u + 3]; }
    }
    for (; b < nrows; ++b) s0 += __ldcg(q + (size_t)b * stride);
    return (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------
// coordinate maps -- exptool/basis/compatibility.py:16-99
// ---------------------------------------------------------------------------
__device__ __forceinline__ double bfe_r_to_xi(double r, int cmap, double inv_scale) {
    // compatibility.py:30-43 ; negatives map to 0.  r/scale is a multiply by the host-rounded reciprocal.
    double out;
    if (cmap == 1) {
        double q = r * inv_scale;
        out = (q - 1.0) / (q + 1.0);
    } else if (cmap == 2) {
        out = log(r);
    } else {
        out = r;
    }
    return (r < 0.0) ? 0.0 : out;
}

__device__ __forceinline__ double bfe_d_xi_to_r(double xi, int cmap, double inv_scale) {
    // compatibility.py:73-80
    if (cmap == 1) return 0.5 * (1.0 - xi) * (1.0 - xi) * inv_scale;
    if (cmap == 2) return exp(-xi);
    return 1.0;
}

__device__ __forceinline__ double bfe_z_to_y(double z, double inv_hscale) {
    // compatibility.py:91 (epsilon 1e-8: the live Python value, not accumulate.c's 1e-10)
    double az = fabs(z);
    double u = fabs(z * inv_hscale);
    // asinh(u), u >= 0.  Only (y - ymin)/dy with y - ymin = O(1..10) is ever used, so ABSOLUTE accuracy ~1e-16 is
    // what matters: log(u + sqrt(u^2+1)) delivers it without asinh()'s small-argument branches (about half the cost).
    double ash = (u < 1.0e150) ? log(u + sqrt(fma(u, u, 1.0))) : asinh(u);
    return (z / (az + 1.0e-8)) * ash;
}

// ---------------------------------------------------------------------------
// EOF bins + bilinear weights -- eof.py:394-422, 443-451
// ---------------------------------------------------------------------------
struct EofBin {
    int node;              // ix*(numy+1)+iy ; the other corners are +1, +ny1, +ny1+1
    int cell;              // ix*numy+iy
    double c00, c10, c01, c11;
};

__device__ __forceinline__ EofBin bfe_eof_bin(const EofGeom& g, double r, double z) {
    // (xi - xmin)/dx as a multiply by the host-rounded reciprocal: <= 1 ulp from the reference's
    // division, i.e. ~1e-14 relative in the bin fractions (X ~ 1e2), far inside the 1e-10 gate
    double X = (bfe_r_to_xi(r, g.cmap, g.inv_ascale) - g.xmin) * g.inv_dx;
    double Y = (bfe_z_to_y(z, g.inv_hscale) - g.ymin) * g.inv_dy;
    int ix = (int)X;                       // truncation, eof.py:404 (NaN -> 0, +-inf saturate)
    int iy = (int)Y;
    if (ix < 0) ix = 0;                    // 410
    if (X < 0.0) X = 0.0;                  // 412
    if (ix >= g.numx) ix = g.numx - 1;     // 414 (X deliberately NOT clamped: 415 is a no-op)
    if (iy < 0) iy = 0;
    if (Y < 0.0) Y = 0.0;
    if (iy >= g.numy) iy = g.numy - 1;
    double delx0 = (double)ix + 1.0 - X;
    double dely0 = (double)iy + 1.0 - Y;
    double delx1 = X - (double)ix;
    double dely1 = Y - (double)iy;
    EofBin b;
    b.node = ix * g.ny1 + iy;
    b.cell = ix * g.numy + iy;
    b.c00 = delx0 * dely0;
    b.c10 = delx1 * dely0;
    b.c01 = delx0 * dely1;
    b.c11 = delx1 * dely1;
    return b;
}

// Cell id only (no weights), for the histogram pass of the cell sort: the index arithmetic is done in FP32 and
// accepted only when the result is provably the FP64 one -- X and Y further than a rigorous error bound from
// every integer (cell edge) and finite.  Otherwise (about 1 particle in 10^3, plus NaN / overflow / denormal
// inputs) the caller recomputes with bfe_eof_bin.  The FP32 chain uses the hardware approximations (sqrt.approx,
// div.approx = __fdividef, lg2.approx = __logf: one MUFU each instead of ~10-20 instructions of IEEE fix-up; the
// histogram kernel was issue-bound on them, ncu profiles/r02_ncu_full_step_kernels.csv).  Error budget: inputs 6e-8
// relative; sqrt.approx 1 ulp, __fdividef 2 ulp (operands here are far inside 2^+-126), __logf 2^-21.4 = 3.6e-7 ABSOLUTE
// for arguments in [0.5, 2] and 3 ulp outside -- its argument u + sqrt(u^2+1) is >= 1, and log r (cmap 2) of r in
// [0.5, 2] has |xi| < 0.7.  Chains: xi (cmap 1) = (q-1)/(q+1): <= 6 ulp; y = sign * log(u + sqrt(u^2+1)): <= 3.6e-7
// absolute + 8 ulp relative (with the z / (|z| + 1e-8) factor).  So |xi_f - xi| and |y_f - y| <= 8e-7 max(1, |.|); the bound below carries a factor 2
// on top and the FP32 rounding of the final scale-and-shift.
__device__ __forceinline__ float bfe_sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ bool bfe_eof_cell_fast(const EofGeom& g, double px, double py, double pz, int& cell) {
    const float xf = (float)px, yf = (float)py, zf = (float)pz;
    const float r = bfe_sqrt_approx(fmaf(xf, xf, fmaf(yf, yf, 1.e-10f)));
    float xi;
    if (g.cmap == 1) {
        const float q = r * (float)g.inv_ascale;
        xi = __fdividef(q - 1.0f, q + 1.0f);
    } else if (g.cmap == 2) {
        xi = __logf(r);
    } else {
        xi = r;
    }
    const float az = fabsf(zf);
    const float u = az * fabsf((float)g.inv_hscale);
    const float ash = __logf(u + bfe_sqrt_approx(fmaf(u, u, 1.0f)));
    const float yy = __fdividef(zf, az + 1.0e-8f) * ash;
    const float idx = (float)g.inv_dx, idy = (float)g.inv_dy;
    const float X = (xi - (float)g.xmin) * idx;
    const float Y = (yy - (float)g.ymin) * idy;
    const float tolx = 1.6e-6f * fabsf(idx) * fmaxf(1.0f, fmaxf(fabsf(xi), fabsf((float)g.xmin))) + 4.0e-7f * fabsf(X) + 1.0e-6f;
    const float toly = 1.6e-6f * fabsf(idy) * fmaxf(1.0f, fmaxf(fabsf(yy), fabsf((float)g.ymin))) + 4.0e-7f * fabsf(Y) + 1.0e-6f;
    // finite, inside float's comfortable range, and away from every cell edge by more than the bound
    bool ok = (fabsf(X) < 1.0e9f) && (fabsf(Y) < 1.0e9f) && (r > 1.0e-18f) && (r < 1.0e18f) && (az < 1.0e18f);
    int ix, iy;
    if (X < -tolx) ix = 0;
    else if (X > (float)g.numx + tolx) ix = g.numx - 1;
    else { ix = (int)X; ok = ok && (fabsf(X - rintf(X)) > tolx); if (ix >= g.numx) ix = g.numx - 1; if (ix < 0) ix = 0; }
    if (Y < -toly) iy = 0;
    else if (Y > (float)g.numy + toly) iy = g.numy - 1;
    else { iy = (int)Y; ok = ok && (fabsf(Y - rintf(Y)) > toly); if (iy >= g.numy) iy = g.numy - 1; if (iy < 0) iy = 0; }
    cell = ix * g.numy + iy;
    return ok;      // NaN anywhere makes a comparison above false -> not ok
}

// cos(phi), sin(phi) of phi = atan2(y, x) without the atan2 (eof.py:532, spheresl.py:613).
__device__ __forceinline__ void bfe_cossin_phi(double x, double y, double& c, double& s) {
    double h2 = x * x + y * y;
    if (h2 > 0.0 && h2 < 1.0e300) {
        double ih = rsqrt(h2);
        // one Newton step: rsqrt() is not correctly rounded
        ih = ih * (1.5 - 0.5 * h2 * ih * ih);
        c = x * ih;
        s = y * ih;
    } else {
        double phi = atan2(y, x);         // zeros / denormals / huge: defer to libm semantics
        sincos(phi, &s, &c);
    }
}

// ---------------------------------------------------------------------------
// EOF field from contracted grids.  G[node][m][6] = (pc, ps, rc, rs, zc, zs) with
// pc[m] = sum_n cos[m,n] potC[m,n,node] etc.  Restates eof.py:842-864 / 1092-1134 with
// the sum over n done once per coefficient set (exact algebra; FP64 order differs).
// ---------------------------------------------------------------------------
struct EofField {
    double p0;   // m = 0 potential
    double p;    // sum over m >= 1
    double fr;   // includes m = 0
    double fp;
    double fz;   // includes m = 0
};

template <int MCAP>
__device__ __forceinline__ EofField bfe_eof_eval(const EofGeom& g, const double* __restrict__ G, int gstride,
                                                 const EofBin& b, double c1, double s1) {
    const double2* n00 = reinterpret_cast<const double2*>(G + (size_t)b.node * gstride);
    const double2* n01 = reinterpret_cast<const double2*>(G + (size_t)(b.node + 1) * gstride);
    const double2* n10 = reinterpret_cast<const double2*>(G + (size_t)(b.node + g.ny1) * gstride);
    const double2* n11 = reinterpret_cast<const double2*>(G + (size_t)(b.node + g.ny1 + 1) * gstride);
    EofField f;
    f.p0 = 0.0; f.p = 0.0; f.fr = 0.0; f.fp = 0.0; f.fz = 0.0;
    double cm = 1.0, sm = 0.0;
#pragma unroll
    for (int m = 0; m <= MCAP; ++m) {
        if (m <= g.mmax) {
            double2 a0 = __ldg(n00 + 3 * m), a1 = __ldg(n00 + 3 * m + 1), a2 = __ldg(n00 + 3 * m + 2);
            double2 b0 = __ldg(n10 + 3 * m), b1 = __ldg(n10 + 3 * m + 1), b2 = __ldg(n10 + 3 * m + 2);
            double2 c0 = __ldg(n01 + 3 * m), c1v = __ldg(n01 + 3 * m + 1), c2 = __ldg(n01 + 3 * m + 2);
            double2 d0 = __ldg(n11 + 3 * m), d1 = __ldg(n11 + 3 * m + 1), d2 = __ldg(n11 + 3 * m + 2);
            double vpc = a0.x * b.c00 + b0.x * b.c10 + c0.x * b.c01 + d0.x * b.c11;
            double vps = a0.y * b.c00 + b0.y * b.c10 + c0.y * b.c01 + d0.y * b.c11;
            double vrc = a1.x * b.c00 + b1.x * b.c10 + c1v.x * b.c01 + d1.x * b.c11;
            double vrs = a1.y * b.c00 + b1.y * b.c10 + c1v.y * b.c01 + d1.y * b.c11;
            double vzc = a2.x * b.c00 + b2.x * b.c10 + c2.x * b.c01 + d2.x * b.c11;
            double vzs = a2.y * b.c00 + b2.y * b.c10 + c2.y * b.c01 + d2.y * b.c11;
            if (m == 0) {
                f.p0 = vpc;
                f.fr = vrc;
                f.fz = vzc;
            } else {
                f.p  += cm * vpc + sm * vps;
                f.fr += cm * vrc + sm * vrs;
                f.fz += cm * vzc + sm * vzs;
                f.fp += (double)m * (sm * vpc - cm * vps);
            }
            // advance to (m+1) phi
            double cn = cm * c1 - sm * s1;
            double sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
    }
    return f;
}

// ---------------------------------------------------------------------------
// Warp-cooperative ("staged") evaluation.
//
// With one point per lane and scattered positions, every LDG.128 of a table row touches 32 different
// 128-byte lines and costs ~32 L1 tag cycles; 84 of them per point bound the direct kernels (ncu,
// profiles/).  Here the rows a point needs are stored as ONE contiguous block per cell (EOF:
// G4[cell][m][corner][3 x double2]; SL: A3[j][(m,l)][3 nodes]), the WARP copies the 32 blocks of its
// 32 points with coalesced 16-byte loads (each instruction covers 512 contiguous-ish bytes: ~5 lines
// instead of 32) into a shared-memory stage, and each lane then reads its own block back conflict-free.
// All 32 lanes must call these functions together (pass a clamped, valid index for idle lanes).
// ---------------------------------------------------------------------------
#define BFE_STAGE_DOUBLE2 (32 * 21 + 8)      // per-warp stage: 32 points x up to 21 chunks (stride kept odd)

template <int NCH>
__device__ __forceinline__ void bfe_warp_stage(const double2* __restrict__ table, size_t blk_chunks, int first_chunk,
                                               int my_block, double2* __restrict__ st, int lane) {
    constexpr int STRIDE = NCH | 1;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < NCH; ++t) {
        const int c = lane + 32 * t;
        const int owner = c / NCH, piece = c - owner * NCH;
        const int oblk = __shfl_sync(0xffffffffu, my_block, owner);
        st[owner * STRIDE + piece] = __ldg(table + (size_t)oblk * blk_chunks + first_chunk + piece);
    }
    __syncwarp();
}

template <int MCAP>
__device__ __forceinline__ EofField bfe_eof_eval_staged(const EofGeom& g, const double2* __restrict__ G4,
                                                        const EofBin& b, double c1, double s1,
                                                        double2* __restrict__ st, int lane) {
    EofField f;
    f.p0 = 0.0; f.p = 0.0; f.fr = 0.0; f.fp = 0.0; f.fz = 0.0;
    double cm = 1.0, sm = 0.0;
    const size_t blk = (size_t)12 * (g.mmax + 1);
    const double2* my = st + lane * 13;
#pragma unroll
    for (int m = 0; m <= MCAP; ++m) {
        if (m <= g.mmax) {
            bfe_warp_stage<12>(G4, blk, 12 * m, b.cell, st, lane);
            // corners 00, 10, 01, 11; three double2 each: (pc,ps), (rc,rs), (zc,zs)
            const double2 a0 = my[0], a1 = my[1], a2 = my[2];
            const double2 b0 = my[3], b1 = my[4], b2 = my[5];
            const double2 c0 = my[6], c1v = my[7], c2 = my[8];
            const double2 d0 = my[9], d1 = my[10], d2 = my[11];
            double vpc = a0.x * b.c00 + b0.x * b.c10 + c0.x * b.c01 + d0.x * b.c11;
            double vps = a0.y * b.c00 + b0.y * b.c10 + c0.y * b.c01 + d0.y * b.c11;
            double vrc = a1.x * b.c00 + b1.x * b.c10 + c1v.x * b.c01 + d1.x * b.c11;
            double vrs = a1.y * b.c00 + b1.y * b.c10 + c1v.y * b.c01 + d1.y * b.c11;
            double vzc = a2.x * b.c00 + b2.x * b.c10 + c2.x * b.c01 + d2.x * b.c11;
            double vzs = a2.y * b.c00 + b2.y * b.c10 + c2.y * b.c01 + d2.y * b.c11;
            if (m == 0) {
                f.p0 = vpc;
                f.fr = vrc;
                f.fz = vzc;
            } else {
                f.p  += cm * vpc + sm * vps;
                f.fr += cm * vrc + sm * vrs;
                f.fz += cm * vzc + sm * vzs;
                f.fp += (double)m * (sm * vpc - cm * vps);
            }
            double cn = cm * c1 - sm * s1;
            double sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
    }
    return f;
}

// ---------------------------------------------------------------------------
// associated Legendre functions -- spheresl.py:664-700 (legendre_R), 706-770 (dlegendre_R)
// ---------------------------------------------------------------------------
template <int LCAP>
struct LegTable {
    double p[LCAP + 1][LCAP + 1];
};

// The recurrences are evaluated with explicitly rounded (non-fused) operations in the
// reference's own operation order: dlegendre divides a cancelling difference by
// (x^2 - 1), so near the poles a 1-ulp change in P is amplified by up to 1e8 in the
// reference itself; matching NumPy's rounding step by step keeps parity there.
#define BFE_MUL(a, b) __dmul_rn((a), (b))
#define BFE_ADD(a, b) __dadd_rn((a), (b))
#define BFE_SUB(a, b) __dadd_rn((a), -(b))
#define BFE_DIV(a, b) __ddiv_rn((a), (b))

// a / k for a small positive integer k that is a compile-time constant after unrolling, with the SAME result as
// __ddiv_rn(a, k) at a fraction of its cost: powers of two are exact scalings; for the odd part d, with
// y = RN(1/d):  q = RN(a y),  r = a - d q (exact in one FMA),  RN(q + r y) is the correctly rounded quotient
// (Markstein's theorem; checked against a/k on 4.2e5 random operands for d = 3..15, tests/golden notes).
__device__ __forceinline__ double bfe_div_int(double a, int k) {
    double scale = 1.0;
    int d = k;
    while ((d & 1) == 0 && d > 1) { d >>= 1; scale *= 0.5; }
    if (d == 1) return a * scale;
    const double dd = (double)d, y = 1.0 / dd;
    const double q = a * y;
    const double r = fma(-dd, q, a);
    return fma(r, y, q) * scale;
}

template <int LCAP>
__device__ __forceinline__ void bfe_legendre(int lmax, double x, LegTable<LCAP>& T) {
    T.p[0][0] = 1.0;
    double pll = 1.0;
    double somx2 = sqrt(BFE_MUL(BFE_SUB(1.0, x), BFE_ADD(1.0, x)));       // spheresl.py:679
    double fact = 1.0;
#pragma unroll
    for (int m = 1; m <= LCAP; ++m) {
        if (m <= lmax) {
            pll = BFE_MUL(pll, BFE_MUL(-fact, somx2));                    // 682
            T.p[m][m] = pll;
            fact += 2.0;
        }
    }
#pragma unroll
    for (int m = 0; m < LCAP; ++m) {
        if (m < lmax) {
            double pl2 = T.p[m][m];
            double pl1 = BFE_MUL(BFE_MUL(x, 2.0 * m + 1.0), pl2);         // 689
            T.p[m + 1][m] = pl1;
#pragma unroll
            for (int l = m + 2; l <= LCAP; ++l) {
                if (l <= lmax) {
                    // (x*(2*l-1)*pl1-(l+m-1)*pl2)/(l-m)                   // 693
                    double a = BFE_MUL(BFE_MUL(x, (double)(2 * l - 1)), pl1);
                    double b = BFE_MUL((double)(l + m - 1), pl2);
                    double v = bfe_div_int(BFE_SUB(a, b), l - m);
                    T.p[l][m] = v;
                    pl2 = pl1;
                    pl1 = v;
                }
            }
        }
    }
}

// Same table with fused multiply-adds and reciprocal constants (a few ulp from the reference): used by
// the accumulation kernels, which need no derivative and therefore no bit-matching.
template <int LCAP>
__device__ __forceinline__ void bfe_legendre_fast(int lmax, double x, LegTable<LCAP>& T) {
    T.p[0][0] = 1.0;
    double pll = 1.0;
    const double somx2 = sqrt((1.0 - x) * (1.0 + x));
    double fact = 1.0;
#pragma unroll
    for (int m = 1; m <= LCAP; ++m) {
        if (m <= lmax) {
            pll *= -fact * somx2;
            T.p[m][m] = pll;
            fact += 2.0;
        }
    }
#pragma unroll
    for (int m = 0; m < LCAP; ++m) {
        if (m < lmax) {
            double pl2 = T.p[m][m];
            double pl1 = x * (2.0 * m + 1.0) * pl2;
            T.p[m + 1][m] = pl1;
#pragma unroll
            for (int l = m + 2; l <= LCAP; ++l) {
                if (l <= lmax) {
                    double v = (x * (double)(2 * l - 1) * pl1 - (double)(l + m - 1) * pl2) * (1.0 / (double)(l - m));
                    T.p[l][m] = v;
                    pl2 = pl1;
                    pl1 = v;
                }
            }
        }
    }
}

// derivative table, spheresl.py:751-768; p must already hold legendre(x)
template <int LCAP>
__device__ __forceinline__ void bfe_dlegendre(int lmax, double x, const LegTable<LCAP>& T, LegTable<LCAP>& D) {
    const double MINEPS = 1.0e-8;
    if (1.0 - fabs(x) < MINEPS) x = (x > 0.0) ? (1.0 - MINEPS) : -(1.0 - MINEPS);
    double somx2 = BFE_DIV(1.0, BFE_SUB(BFE_MUL(x, x), 1.0));             // 757
    D.p[0][0] = 0.0;
#pragma unroll
    for (int l = 1; l <= LCAP; ++l) {
        if (l <= lmax) {
#pragma unroll
            for (int m = 0; m < l; ++m) {
                // somx2*(x*l*p[l][m] - (l+m)*p[l-1][m])                   // 763
                double a = BFE_MUL(BFE_MUL(x, (double)l), T.p[l][m]);
                double b = BFE_MUL((double)(l + m), T.p[l - 1][m]);
                D.p[l][m] = BFE_MUL(somx2, BFE_SUB(a, b));
            }
            D.p[l][l] = BFE_MUL(BFE_MUL(BFE_MUL(somx2, x), (double)l), T.p[l][l]);   // 765
        }
    }
}

// ---------------------------------------------------------------------------
// SL radial bin -- spheresl.py:123-142 / 309-328
// ---------------------------------------------------------------------------
struct SlBin {
    int i;          // lower node of the interpolation interval, clamped to [0, numr-2]
    double x1, x2;  // linear weights of nodes i, i+1
    double fac;     // d_xi_to_r(xi)/dxi
};

__device__ __forceinline__ SlBin bfe_sl_bin(const SlGeom& g, const double* __restrict__ xi, double r) {
    double x = bfe_r_to_xi(r, g.cmap, g.inv_scale);
    if (g.cmap == 1) {
        if (x < -1.0) x = -1.0;
        if (x >= 1.0) x = 1.0 - 1.0e-08;
    }
    double fi = floor((x - g.xi0) * g.inv_dxi);
    int i;
    if (!(fi >= 0.0)) i = 0;
    else if (fi > (double)(g.numr - 2)) i = g.numr - 2;
    else i = (int)fi;
    SlBin b;
    b.i = i;
    b.x1 = (__ldg(xi + i + 1) - x) * g.inv_dxi;
    b.x2 = (x - __ldg(xi + i)) * g.inv_dxi;
    b.fac = bfe_d_xi_to_r(x, g.cmap, g.inv_scale) * g.inv_dxi;
    return b;
}

// ---------------------------------------------------------------------------
// SL field from contracted rows.  A[i][q] (double2, q = l(l+1)/2 + m) holds
//   .x = sum_n c[l^2 + 2m-1 (or l^2 for m=0), n] eftable[l,n,i]/sqrt(ev[l,n])   (cosine row)
//   .y = sum_n c[l^2 + 2m, n] ...                                              (sine row; 0 for m=0)
// Restates spheresl.py:1041-1086 / 1200-1228 / 1289-1339 with the sum over n done once per
// coefficient set.  The loops run m-outer / l-inner so that only two Legendre values of the
// current column are live (dlegendre needs P_l^m and P_{l-1}^m of the SAME m, spheresl.py:763):
// every P_l^m, dP_l^m is produced by exactly the reference's chain of rounded operations.
// ---------------------------------------------------------------------------
struct SlField {
    double pot0, pot1, potr, pott, potp;
};

template <int LCAP>
__device__ __forceinline__ SlField bfe_sl_eval(const SlGeom& g, const double2* __restrict__ A, int qstride,
                                               const double* __restrict__ p0tab, const double* __restrict__ fac,
                                               const SlBin& b, double costh, double c1, double s1,
                                               bool trig_index_l) {
    // nodes: potential uses (i, i+1); derivative stencil uses (j-1, j, j+1), j = max(i,1)
    const int j = (b.i == 0) ? 1 : b.i;
    const double2* rowm = A + (size_t)(j - 1) * qstride;
    const double2* row0 = A + (size_t)j * qstride;
    const double2* rowp = A + (size_t)(j + 1) * qstride;
    const double pm = __ldg(p0tab + j - 1), pc = __ldg(p0tab + j), pp = __ldg(p0tab + j + 1);
    // weights so that pot = (wA*A[j-1] + wB*A[j] + wC*A[j+1]) * P0
    double P0, wA, wB, wC;
    if (b.i == 0) { P0 = b.x1 * pm + b.x2 * pc; wA = b.x1; wB = b.x2; wC = 0.0; }
    else          { P0 = b.x1 * pc + b.x2 * pp; wA = 0.0; wB = b.x1; wC = b.x2; }
    wA *= P0; wB *= P0; wC *= P0;
    const double dA = b.fac * (b.x2 - 0.5) * pm, dB = b.fac * (-2.0 * b.x2) * pc, dC = b.fac * (b.x2 + 0.5) * pp;

    const double x = costh;
    const double somx2 = sqrt(BFE_MUL(BFE_SUB(1.0, x), BFE_ADD(1.0, x)));     // spheresl.py:679
    double xd = x;                                                            // 751-755
    if (1.0 - fabs(xd) < 1.0e-8) xd = (xd > 0.0) ? (1.0 - 1.0e-8) : -(1.0 - 1.0e-8);
    const double dsom = BFE_DIV(1.0, BFE_SUB(BFE_MUL(xd, xd), 1.0));          // 757

    SlField f;
    f.pot0 = 0.0; f.pot1 = 0.0; f.potr = 0.0; f.pott = 0.0; f.potp = 0.0;
    double pmm = 1.0, fact = 1.0;
    double cm = 1.0, sm = 0.0;
#pragma unroll
    for (int m = 0; m <= LCAP; ++m) {
        if (m <= g.lmax) {
            if (m > 0) {
                pmm = BFE_MUL(pmm, BFE_MUL(-fact, somx2));                    // 682
                fact += 2.0;
                double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1;
                cm = cn; sm = sn;
            }
            double pl2 = 0.0, pl1 = pmm;                                       // P_{l-1}^m, P_l^m at l = m
            double cl = cm, sl = sm;                                           // cos/sin(l phi) at l = m
#pragma unroll
            for (int l = m; l <= LCAP; ++l) {
                if (l <= g.lmax) {
                    double P, dP;
                    if (l == m) {
                        P = pmm;
                        dP = (l == 0) ? 0.0 : BFE_MUL(BFE_MUL(BFE_MUL(dsom, xd), (double)l), P);          // 765
                    } else {
                        if (l == m + 1) P = BFE_MUL(BFE_MUL(x, 2.0 * m + 1.0), pl1);                      // 689
                        else P = bfe_div_int(BFE_SUB(BFE_MUL(BFE_MUL(x, (double)(2 * l - 1)), pl1),
                                                     BFE_MUL((double)(l + m - 1), pl2)), l - m);         // 693
                        dP = BFE_MUL(dsom, BFE_SUB(BFE_MUL(BFE_MUL(xd, (double)l), P),
                                                   BFE_MUL((double)(l + m), pl1)));                       // 763
                        pl2 = pl1; pl1 = P;
                        double cn = cl * c1 - sl * s1, sn = sl * c1 + cl * s1;
                        cl = cn; sl = sn;
                    }
                    const int q = (l * (l + 1)) / 2 + m;
                    const double2 am = __ldg(rowm + q), a0 = __ldg(row0 + q), ap = __ldg(rowp + q);
                    const double fl = __ldg(fac + l * (g.lmax + 1) + m);
                    const double spc = wA * am.x + wB * a0.x + wC * ap.x;
                    const double sdc = dA * am.x + dB * a0.x + dC * ap.x;
                    if (m == 0) {
                        if (l == 0) {
                            f.pot0 = fl * spc;
                            f.potr += fl * sdc;
                        } else {
                            f.pot1 += fl * P * spc;
                            f.potr += fl * P * sdc;
                            f.pott += fl * dP * spc;
                        }
                    } else {
                        const double sps = wA * am.y + wB * a0.y + wC * ap.y;
                        const double sds = dA * am.y + dB * a0.y + dC * ap.y;
                        const double ct = trig_index_l ? cl : cm;
                        const double st = trig_index_l ? sl : sm;
                        const double Ap = spc * ct + sps * st;
                        const double Ad = sdc * ct + sds * st;
                        const double Bp = sps * ct - spc * st;
                        const double fP = fl * P;
                        f.pot1 += fP * Ap;
                        f.potr += fP * Ad;
                        f.pott += fl * dP * Ap;
                        f.potp += fP * (double)m * Bp;
                    }
                }
            }
        }
    }
    return f;
}

// factorial_return factors fac[l*(lmax+1)+m] (spheresl.py:26-35) either through a device pointer (__ldg) or BY VALUE in
// the kernel parameter block (SlFacP): with the loops unrolled the indices are constants and the factor becomes a
// constant-bank operand of the multiply -- 28 fewer load instructions per point at lmax = 6, each of which cost a full
// 8-wavefront pass of the L1 data pipe (ncu l1tex__data_pipe_lsu_wavefronts is what bounds the coherent per-lane kernels).
struct SlFacP { double v[49]; };               // (lmax+1)^2 <= 49: the block kernels cover lmax 4 and 6
inline SlFacP bfe_sl_facp(const bfe_sl* h) {     // valid for lmax <= 6 (the block kernels)
    SlFacP f;
    const int nf = (h->g.lmax + 1) * (h->g.lmax + 1);
    for (int k = 0; k < 49; ++k) f.v[k] = k < nf ? h->fac_host[k] : 0.0;
    return f;
}
__device__ __forceinline__ double bfe_fac_at(const double* fac, int k) { return __ldg(fac + k); }
__device__ __forceinline__ double bfe_fac_at(const SlFacP& fac, int k) { return fac.v[k]; }

// interval blocks of A3 are padded to an even number of double2 (32-byte alignment for the 256-bit loads below)
#define BFE_A3_STRIDE(npair) ((3 * (npair) + 1) & ~1)

// Staged SL evaluation; valid only when g.lmax == LCAP (compile-time chunk counts).
// A3[j][chunk], chunk = 3*(offm(m) + l - m) + node, node 0,1,2 = rows j-1, j, j+1; offm(m) = sum_{m'<m}(LCAP-m'+1).
template <int LCAP>
__device__ __forceinline__ SlField bfe_sl_eval_staged(const SlGeom& g, const double2* __restrict__ A3,
                                                      const double* __restrict__ p0tab, const double* __restrict__ fac,
                                                      const SlBin& b, double costh, double c1, double s1,
                                                      bool trig_index_l, double2* __restrict__ st, int lane) {
    constexpr int NPAIR = (LCAP + 1) * (LCAP + 2) / 2;
    const int j = (b.i == 0) ? 1 : b.i;
    const double pm = __ldg(p0tab + j - 1), pc = __ldg(p0tab + j), pp = __ldg(p0tab + j + 1);
    double P0, wA, wB, wC;
    if (b.i == 0) { P0 = b.x1 * pm + b.x2 * pc; wA = b.x1; wB = b.x2; wC = 0.0; }
    else          { P0 = b.x1 * pc + b.x2 * pp; wA = 0.0; wB = b.x1; wC = b.x2; }
    wA *= P0; wB *= P0; wC *= P0;
    const double dA = b.fac * (b.x2 - 0.5) * pm, dB = b.fac * (-2.0 * b.x2) * pc, dC = b.fac * (b.x2 + 0.5) * pp;

    const double x = costh;
    const double somx2 = sqrt(BFE_MUL(BFE_SUB(1.0, x), BFE_ADD(1.0, x)));
    double xd = x;
    if (1.0 - fabs(xd) < 1.0e-8) xd = (xd > 0.0) ? (1.0 - 1.0e-8) : -(1.0 - 1.0e-8);
    const double dsom = BFE_DIV(1.0, BFE_SUB(BFE_MUL(xd, xd), 1.0));

    SlField f;
    f.pot0 = 0.0; f.pot1 = 0.0; f.potr = 0.0; f.pott = 0.0; f.potp = 0.0;
    double pmm = 1.0, fact = 1.0;
    double cm = 1.0, sm = 0.0;
    int offm = 0;
#pragma unroll
    for (int m = 0; m <= LCAP; ++m) {
        constexpr int dummy = 0; (void)dummy;
        const int nl = LCAP - m + 1;                       // l values in this column (compile time after unrolling)
        // stage this column: 3*nl chunks per point
        {
            const int NCHr = 3 * nl, STRIDE = NCHr | 1;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 3 * (LCAP + 1); ++t) {
                if (t < NCHr) {
                    const int c = lane + 32 * t;
                    const int owner = c / NCHr, piece = c - owner * NCHr;
                    const int oj = __shfl_sync(0xffffffffu, j, owner);
                    st[owner * STRIDE + piece] = __ldg(A3 + (size_t)oj * BFE_A3_STRIDE(NPAIR) + 3 * offm + piece);
                }
            }
            __syncwarp();
        }
        const double2* my = st + lane * ((3 * nl) | 1);
        if (m > 0) {
            pmm = BFE_MUL(pmm, BFE_MUL(-fact, somx2));
            fact += 2.0;
            double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
        double pl2 = 0.0, pl1 = pmm;
        double cl = cm, sl = sm;
#pragma unroll
        for (int l = m; l <= LCAP; ++l) {
            double P, dP;
            if (l == m) {
                P = pmm;
                dP = (l == 0) ? 0.0 : BFE_MUL(BFE_MUL(BFE_MUL(dsom, xd), (double)l), P);
            } else {
                if (l == m + 1) P = BFE_MUL(BFE_MUL(x, 2.0 * m + 1.0), pl1);
                else P = bfe_div_int(BFE_SUB(BFE_MUL(BFE_MUL(x, (double)(2 * l - 1)), pl1),
                                             BFE_MUL((double)(l + m - 1), pl2)), l - m);
                dP = BFE_MUL(dsom, BFE_SUB(BFE_MUL(BFE_MUL(xd, (double)l), P), BFE_MUL((double)(l + m), pl1)));
                pl2 = pl1; pl1 = P;
                double cn = cl * c1 - sl * s1, sn = sl * c1 + cl * s1;
                cl = cn; sl = sn;
            }
            const double2 am = my[3 * (l - m)], a0 = my[3 * (l - m) + 1], ap = my[3 * (l - m) + 2];
            const double fl = __ldg(fac + l * (LCAP + 1) + m);
            const double spc = wA * am.x + wB * a0.x + wC * ap.x;
            const double sdc = dA * am.x + dB * a0.x + dC * ap.x;
            if (m == 0) {
                if (l == 0) {
                    f.pot0 = fl * spc;
                    f.potr += fl * sdc;
                } else {
                    f.pot1 += fl * P * spc;
                    f.potr += fl * P * sdc;
                    f.pott += fl * dP * spc;
                }
            } else {
                const double sps = wA * am.y + wB * a0.y + wC * ap.y;
                const double sds = dA * am.y + dB * a0.y + dC * ap.y;
                const double ct = trig_index_l ? cl : cm;
                const double stt = trig_index_l ? sl : sm;
                const double Ap = spc * ct + sps * stt;
                const double Ad = sdc * ct + sds * stt;
                const double Bp = sps * ct - spc * stt;
                const double fP = fl * P;
                f.pot1 += fP * Ap;
                f.potr += fP * Ad;
                f.pott += fl * dP * Ap;
                f.potp += fP * (double)m * Bp;
            }
        }
        offm += nl;
    }
    return f;
}

// ---------------------------------------------------------------------------
// Per-lane evaluation from the per-cell / per-interval BLOCKS (G4, A3) with 256-bit loads.
//
// The per-lane kernels above are bound by the L1 tag stage: every divergent load instruction costs ~32 tag
// cycles whatever its width (ncu, profiles/).  G4[cell] (EOF: 192 B per harmonic) and A3[j] (SL: one 48-B entry
// per (m,l), consecutive in the evaluation order) are contiguous and 32-byte aligned, so the same operands come in
// with HALF the load instructions as 32-byte (LDG.E.256) loads; all of a point's loads are independent of each
// other (addresses = block base + constant), so they are in flight together.  Arithmetic is unchanged.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void bfe_ldg256(const double2* p, double2& a, double2& b) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}

// How a block evaluator fetches 32 bytes: LdGlobal256 = one 256-bit read-only global load (the caller-order and
// key-ordered per-lane kernels); LdGeneric = two 128-bit GENERIC loads, for a base pointer that is either a block staged
// in shared memory by a TMA bulk copy or, for the points of a tile whose block was not staged, the block in global memory
// (bfe_orbit_sort.cu).  A warp-uniform 16-byte shared-memory load costs ~1.5 cycles of the SM's data pipe against ~8.3
// for a 32-byte global load that hits L1 (profiles/probes/lds_broadcast_probe.cu): 2.8x less per byte.
struct LdGlobal256 {
    static __device__ __forceinline__ void ld(const double2* p, double2& a, double2& b) { bfe_ldg256(p, a, b); }
};
struct LdGeneric {
    static __device__ __forceinline__ void ld(const double2* p, double2& a, double2& b) { a = p[0]; b = p[1]; }
};

template <int MCAP, class LD = LdGlobal256>
__device__ __forceinline__ EofField bfe_eof_eval_base(const EofGeom& g, const double2* base, const EofBin& b, double c1, double s1);

template <int MCAP>
__device__ __forceinline__ EofField bfe_eof_eval_blk(const EofGeom& g, const double2* __restrict__ G4,
                                                     const EofBin& b, double c1, double s1) {
    return bfe_eof_eval_base<MCAP, LdGlobal256>(g, G4 + (size_t)b.cell * (size_t)(12 * (g.mmax + 1)), b, c1, s1);
}

template <int MCAP, class LD>
__device__ __forceinline__ EofField bfe_eof_eval_base(const EofGeom& g, const double2* base, const EofBin& b, double c1, double s1) {
    EofField f;
    f.p0 = 0.0; f.p = 0.0; f.fr = 0.0; f.fp = 0.0; f.fz = 0.0;
    double cm = 1.0, sm = 0.0;
#pragma unroll
    for (int m = 0; m <= MCAP; ++m) {
        if (m <= g.mmax) {
            // corners 00, 10, 01, 11; three double2 each: (pc,ps), (rc,rs), (zc,zs)
            const double2* q = base + 12 * m;
            double2 a0, a1, a2, b0, b1, b2, c0, c1v, c2, d0, d1, d2;
            LD::ld(q, a0, a1); LD::ld(q + 2, a2, b0); LD::ld(q + 4, b1, b2);
            LD::ld(q + 6, c0, c1v); LD::ld(q + 8, c2, d0); LD::ld(q + 10, d1, d2);
            double vpc = a0.x * b.c00 + b0.x * b.c10 + c0.x * b.c01 + d0.x * b.c11;
            double vps = a0.y * b.c00 + b0.y * b.c10 + c0.y * b.c01 + d0.y * b.c11;
            double vrc = a1.x * b.c00 + b1.x * b.c10 + c1v.x * b.c01 + d1.x * b.c11;
            double vrs = a1.y * b.c00 + b1.y * b.c10 + c1v.y * b.c01 + d1.y * b.c11;
            double vzc = a2.x * b.c00 + b2.x * b.c10 + c2.x * b.c01 + d2.x * b.c11;
            double vzs = a2.y * b.c00 + b2.y * b.c10 + c2.y * b.c01 + d2.y * b.c11;
            if (m == 0) {
                f.p0 = vpc;
                f.fr = vrc;
                f.fz = vzc;
            } else {
                f.p  += cm * vpc + sm * vps;
                f.fr += cm * vrc + sm * vrs;
                f.fz += cm * vzc + sm * vzs;
                f.fp += (double)m * (sm * vpc - cm * vps);
            }
            double cn = cm * c1 - sm * s1;
            double sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
    }
    return f;
}

// ---------------------------------------------------------------------------
// Per-INTERVAL polynomial blocks A4 (FP64; the default per-lane evaluation since round 2).
// The per-lane field kernels are bound by the register-file WRITE port: every FP64 result takes two write cycles of its
// scheduler and every 32 loaded bytes per lane one, and the two add up (profiles/probes/fp64_mix_probe.cu: 2 LDG.256 per
// 32 FP64 instructions lower the FP64 rate from 1.76 to 1.39 per clk, the 64 / 80 the model gives; ncu: the fused kernel's
// time is the sum of its load phase and its FP64 phase, whatever the occupancy).  So the lever is the instruction count.
// On the interval i the radial interpolation of spheresl.py:153-155 is linear in x2:
//     potential part   P0 (x1 a_i + x2 a_{i+1})                         = P0 (a_i + x2 S1),       S1 = a_{i+1} - a_i
//     derivative part  fac ((x2 - 1/2) u_{j-1} - 2 x2 u_j + (x2 + 1/2) u_{j+1}) = fac (D1 + x2 D2),   u_k = p0[k] a_k, j = max(i, 1)
//     D1 = (u_{j+1} - u_{j-1}) / 2,  D2 = u_{j-1} - 2 u_j + u_{j+1}
// with P0 and fac common to all (l, m): they multiply the five sums once at the end.  The block of interval i holds, in the
// evaluation order (m outer, l inner), {a_i, S1, D1, D2} for the cosine row of the m = 0 entries (32 B each) and for the
// cosine and sine rows of the others (64 B each: {a.x, a.y, S1.x, S1.y}, {D1.x, D1.y, D2.x, D2.y}), every value already
// multiplied by the factorial factor of (l, m).  Per (l, m > 0): 4 FMAs of interpolation instead of 12 and no factor
// multiplies; the azimuthal factor m of potp is applied once per column.  1568 B per interval at lmax = 6 (A3: 1344 B).
// ---------------------------------------------------------------------------
#define BFE_A4_BYTES(lmax) (32 * ((lmax) + 1) + 64 * ((((lmax) + 1) * ((lmax) + 2)) / 2 - ((lmax) + 1)))

template <int LCAP>
__device__ __forceinline__ SlField bfe_sl_eval_poly(const SlGeom& g, const void* __restrict__ A4,
                                                    const double* __restrict__ p0tab, const SlBin& b, double costh,
                                                    double c1, double s1, bool trig_index_l) {
    const double2* blk = reinterpret_cast<const double2*>(static_cast<const char*>(A4) + (size_t)b.i * BFE_A4_BYTES(LCAP));
    const double P0 = b.x1 * __ldg(p0tab + b.i) + b.x2 * __ldg(p0tab + b.i + 1);
    const double x2 = b.x2;

    const double x = costh;
    const double somx2 = sqrt(BFE_MUL(BFE_SUB(1.0, x), BFE_ADD(1.0, x)));
    double xd = x;
    if (1.0 - fabs(xd) < 1.0e-8) xd = (xd > 0.0) ? (1.0 - 1.0e-8) : -(1.0 - 1.0e-8);
    const double dsom = BFE_DIV(1.0, BFE_SUB(BFE_MUL(xd, xd), 1.0));

    double q0 = 0.0, q1 = 0.0, qr = 0.0, qt = 0.0, qp = 0.0;
    double pmm = 1.0, fact = 1.0;
    double cm = 1.0, sm = 0.0;
    int off2 = 2 * (LCAP + 1);                     // double2 index of the first m > 0 entry
#pragma unroll
    for (int m = 0; m <= LCAP; ++m) {
        if (m > 0) {
            pmm = BFE_MUL(pmm, BFE_MUL(-fact, somx2));
            fact += 2.0;
            double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
        double pl2 = 0.0, pl1 = pmm;
        double cl = cm, sl = sm;
        double colp = 0.0;
#pragma unroll
        for (int l = m; l <= LCAP; ++l) {
            double P, dP;
            if (l == m) {
                P = pmm;
                dP = (l == 0) ? 0.0 : BFE_MUL(BFE_MUL(BFE_MUL(dsom, xd), (double)l), P);
            } else {
                if (l == m + 1) P = BFE_MUL(BFE_MUL(x, 2.0 * m + 1.0), pl1);
                else P = bfe_div_int(BFE_SUB(BFE_MUL(BFE_MUL(x, (double)(2 * l - 1)), pl1),
                                             BFE_MUL((double)(l + m - 1), pl2)), l - m);
                dP = BFE_MUL(dsom, BFE_SUB(BFE_MUL(BFE_MUL(xd, (double)l), P), BFE_MUL((double)(l + m), pl1)));
                pl2 = pl1; pl1 = P;
                double cn = cl * c1 - sl * s1, sn = sl * c1 + cl * s1;
                cl = cn; sl = sn;
            }
            if (m == 0) {
                double2 e0, e1;                                        // {a, S1}, {D1, D2}
                bfe_ldg256(blk + 2 * l, e0, e1);
                const double spc = e0.x + x2 * e0.y;
                const double sdc = e1.x + x2 * e1.y;
                if (l == 0) {
                    q0 = spc;
                    qr += sdc;
                } else {
                    q1 += P * spc;
                    qr += P * sdc;
                    qt += dP * spc;
                }
            } else {
                double2 a, sd, d1, d2;                                 // {a.x, a.y}, {S1.x, S1.y}, {D1.x, D1.y}, {D2.x, D2.y}
                bfe_ldg256(blk + off2 + 4 * (l - m), a, sd);
                bfe_ldg256(blk + off2 + 4 * (l - m) + 2, d1, d2);
                const double spc = a.x + x2 * sd.x, sps = a.y + x2 * sd.y;
                const double sdc = d1.x + x2 * d2.x, sds = d1.y + x2 * d2.y;
                const double ct = trig_index_l ? cl : cm;
                const double stt = trig_index_l ? sl : sm;
                const double Ap = spc * ct + sps * stt;
                const double Ad = sdc * ct + sds * stt;
                const double Bp = sps * ct - spc * stt;
                q1 += P * Ap;
                qr += P * Ad;
                qt += dP * Ap;
                colp += P * Bp;
            }
        }
        if (m > 0) {
            qp += (double)m * colp;
            off2 += 4 * (LCAP - m + 1);
        }
    }
    SlField f;
    f.pot0 = P0 * q0; f.pot1 = P0 * q1; f.potr = b.fac * qr; f.pott = P0 * qt; f.potp = P0 * qp;
    return f;
}

// valid only when g.lmax == LCAP (LCAP = 6: 84 double2 = 1344 B per interval; LCAP = 4: 45 padded to 46)
template <int LCAP, typename FacT, class LD = LdGlobal256>
__device__ __forceinline__ SlField bfe_sl_eval_base(const SlGeom& g, const double2* base,
                                                    const double* __restrict__ p0tab, const FacT& fac,
                                                    const SlBin& b, double costh, double c1, double s1, bool trig_index_l);

template <int LCAP, typename FacT>
__device__ __forceinline__ SlField bfe_sl_eval_blk(const SlGeom& g, const void* __restrict__ A4,
                                                   const double* __restrict__ p0tab, const FacT& fac,
                                                   const SlBin& b, double costh, double c1, double s1,
                                                   bool trig_index_l) {
    (void)fac;                                   // the factorial factors are folded into the A4 blocks
    return bfe_sl_eval_poly<LCAP>(g, A4, p0tab, b, costh, c1, s1, trig_index_l);
}

// base = the block of radial index j = max(b.i, 1): A3 + j * BFE_A3_STRIDE(NPAIR), or its staged copy
template <int LCAP, typename FacT, class LD>
__device__ __forceinline__ SlField bfe_sl_eval_base(const SlGeom& g, const double2* base,
                                                    const double* __restrict__ p0tab, const FacT& fac,
                                                    const SlBin& b, double costh, double c1, double s1, bool trig_index_l) {
    const int j = (b.i == 0) ? 1 : b.i;
    const double pm = __ldg(p0tab + j - 1), pc = __ldg(p0tab + j), pp = __ldg(p0tab + j + 1);
    double P0, wA, wB, wC;
    if (b.i == 0) { P0 = b.x1 * pm + b.x2 * pc; wA = b.x1; wB = b.x2; wC = 0.0; }
    else          { P0 = b.x1 * pc + b.x2 * pp; wA = 0.0; wB = b.x1; wC = b.x2; }
    wA *= P0; wB *= P0; wC *= P0;
    const double dA = b.fac * (b.x2 - 0.5) * pm, dB = b.fac * (-2.0 * b.x2) * pc, dC = b.fac * (b.x2 + 0.5) * pp;

    const double x = costh;
    const double somx2 = sqrt(BFE_MUL(BFE_SUB(1.0, x), BFE_ADD(1.0, x)));
    double xd = x;
    if (1.0 - fabs(xd) < 1.0e-8) xd = (xd > 0.0) ? (1.0 - 1.0e-8) : -(1.0 - 1.0e-8);
    const double dsom = BFE_DIV(1.0, BFE_SUB(BFE_MUL(xd, xd), 1.0));

    SlField f;
    f.pot0 = 0.0; f.pot1 = 0.0; f.potr = 0.0; f.pott = 0.0; f.potp = 0.0;
    double pmm = 1.0, fact = 1.0;
    double cm = 1.0, sm = 0.0;
    int offm = 0;
#pragma unroll
    for (int m = 0; m <= LCAP; ++m) {
        const int nl = LCAP - m + 1;                       // entries of this column (compile time after unrolling)
        // the column's 3*nl double2 start at double2 index 3*offm of the block: fetch the covering 32-byte pairs
        double2 v[3 * (LCAP + 1) + 1];
        {
            const int first = 3 * offm, count = 3 * nl;
            const int pa = first / 2, pb = (first + count - 1) / 2;
#pragma unroll
            for (int pi = 0; pi < (3 * (LCAP + 1)) / 2 + 1; ++pi) {
                const int pr = pa + pi;
                if (pr <= pb) {
                    double2 lo, hi;
                    LD::ld(base + 2 * pr, lo, hi);
                    if (2 * pr >= first) v[2 * pr - first] = lo;
                    if (2 * pr + 1 < first + count) v[2 * pr + 1 - first] = hi;
                }
            }
        }
        if (m > 0) {
            pmm = BFE_MUL(pmm, BFE_MUL(-fact, somx2));
            fact += 2.0;
            double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
        double pl2 = 0.0, pl1 = pmm;
        double cl = cm, sl = sm;
#pragma unroll
        for (int l = m; l <= LCAP; ++l) {
            double P, dP;
            if (l == m) {
                P = pmm;
                dP = (l == 0) ? 0.0 : BFE_MUL(BFE_MUL(BFE_MUL(dsom, xd), (double)l), P);
            } else {
                if (l == m + 1) P = BFE_MUL(BFE_MUL(x, 2.0 * m + 1.0), pl1);
                else P = bfe_div_int(BFE_SUB(BFE_MUL(BFE_MUL(x, (double)(2 * l - 1)), pl1),
                                             BFE_MUL((double)(l + m - 1), pl2)), l - m);
                dP = BFE_MUL(dsom, BFE_SUB(BFE_MUL(BFE_MUL(xd, (double)l), P), BFE_MUL((double)(l + m), pl1)));
                pl2 = pl1; pl1 = P;
                double cn = cl * c1 - sl * s1, sn = sl * c1 + cl * s1;
                cl = cn; sl = sn;
            }
            const double2 am = v[3 * (l - m)], a0 = v[3 * (l - m) + 1], ap = v[3 * (l - m) + 2];
            const double fl = bfe_fac_at(fac, l * (LCAP + 1) + m);
            const double spc = wA * am.x + wB * a0.x + wC * ap.x;
            const double sdc = dA * am.x + dB * a0.x + dC * ap.x;
            if (m == 0) {
                if (l == 0) {
                    f.pot0 = fl * spc;
                    f.potr += fl * sdc;
                } else {
                    f.pot1 += fl * P * spc;
                    f.potr += fl * P * sdc;
                    f.pott += fl * dP * spc;
                }
            } else {
                const double sps = wA * am.y + wB * a0.y + wC * ap.y;
                const double sds = dA * am.y + dB * a0.y + dC * ap.y;
                const double ct = trig_index_l ? cl : cm;
                const double stt = trig_index_l ? sl : sm;
                const double Ap = spc * ct + sps * stt;
                const double Ad = sdc * ct + sds * stt;
                const double Bp = sps * ct - spc * stt;
                const double fP = fl * P;
                f.pot1 += fP * Ap;
                f.potr += fP * Ad;
                f.pott += fl * dP * Ap;
                f.potp += fP * (double)m * Bp;
            }
        }
        offm += nl;
    }
    return f;
}

// ---------------------------------------------------------------------------
// Fields.return_forces_cart -- potential.py:455-497
// ---------------------------------------------------------------------------
struct CartForce {
    double fxd, fxh, fyd, fyh, fzd, fzh, pd, ph;
};

// ---------------------------------------------------------------------------
// FP32-TABLE mode (option "table_fp32"; BASELINE north_star: "<= 1e-5 where FP32 table interpolation is used").
// The contracted blocks are stored as float (G4f: 24 floats per (cell, m); A3f: 8 floats per (m,l) entry, see below), read
// with the same 256-bit loads -- 8 values each, so HALF the L1 tag cycles of the FP64 blocks, which is what bounds
// these kernels (ncu l1tex 92 %, profiles/r01_ncu_full_blk_kernels.csv).  Only the table VALUES are rounded to
// float (6e-8 relative); bins, weights, Legendre functions, trig factors and all sums stay FP64.
// ---------------------------------------------------------------------------
#define BFE_A3F_STRIDE(npair) (8 * (npair))

__device__ __forceinline__ void bfe_ldg256f(const float* p, float* v) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}

template <int MCAP>
__device__ __forceinline__ EofField bfe_eof_eval_blk32(const EofGeom& g, const float* __restrict__ G4f,
                                                       const EofBin& b, double c1, double s1) {
    const float* base = G4f + (size_t)b.cell * (size_t)(24 * (g.mmax + 1));
    EofField f;
    f.p0 = 0.0; f.p = 0.0; f.fr = 0.0; f.fp = 0.0; f.fz = 0.0;
    double cm = 1.0, sm = 0.0;
#pragma unroll
    for (int m = 0; m <= MCAP; ++m) {
        if (m <= g.mmax) {
            // v[corner*6 + field*2 + cs], corners 00, 10, 01, 11
            float v[24];
            bfe_ldg256f(base + 24 * m, v); bfe_ldg256f(base + 24 * m + 8, v + 8); bfe_ldg256f(base + 24 * m + 16, v + 16);
            double vpc = (double)v[0] * b.c00 + (double)v[6] * b.c10 + (double)v[12] * b.c01 + (double)v[18] * b.c11;
            double vps = (double)v[1] * b.c00 + (double)v[7] * b.c10 + (double)v[13] * b.c01 + (double)v[19] * b.c11;
            double vrc = (double)v[2] * b.c00 + (double)v[8] * b.c10 + (double)v[14] * b.c01 + (double)v[20] * b.c11;
            double vrs = (double)v[3] * b.c00 + (double)v[9] * b.c10 + (double)v[15] * b.c01 + (double)v[21] * b.c11;
            double vzc = (double)v[4] * b.c00 + (double)v[10] * b.c10 + (double)v[16] * b.c01 + (double)v[22] * b.c11;
            double vzs = (double)v[5] * b.c00 + (double)v[11] * b.c10 + (double)v[17] * b.c01 + (double)v[23] * b.c11;
            if (m == 0) {
                f.p0 = vpc;
                f.fr = vrc;
                f.fz = vzc;
            } else {
                f.p  += cm * vpc + sm * vps;
                f.fr += cm * vrc + sm * vrs;
                f.fz += cm * vzc + sm * vzs;
                f.fp += (double)m * (sm * vpc - cm * vps);
            }
            double cn = cm * c1 - sm * s1;
            double sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
    }
    return f;
}

// A3f block of radial interval i (0 <= i <= numr-2), one 32-byte entry per (m,l) in evaluation order:
//   { a_i.x, a_i.y, a_{i+1}.x, a_{i+1}.y, D1.x, D1.y, D2.x, D2.y },   u_k = p0[k] a_k,  j = max(i, 1),
//   D1 = (u_{j+1} - u_{j-1}) / 2,   D2 = u_{j-1} - 2 u_j + u_{j+1}
// so that the derivative stencil of get_halo_dens_pot_force (spheresl.py:153-155)
//   (x2 - 1/2) u_{j-1} - 2 x2 u_j + (x2 + 1/2) u_{j+1}  =  x2 D2 + D1
// is formed from differences taken in FP64 BEFORE the rounding to float: storing the three node values as
// floats and differencing afterwards amplifies their 6e-8 rounding by the ~1e3-1e4 cancellation of the stencil
// (measured 3e-4 in the radial force; profiles/r01_blk_ab.json).
template <int LCAP, typename FacT>
__device__ __forceinline__ SlField bfe_sl_eval_blk32(const SlGeom& g, const float* __restrict__ A3f,
                                                     const double* __restrict__ p0tab, const FacT& fac,
                                                     const SlBin& b, double costh, double c1, double s1,
                                                     bool trig_index_l) {
    constexpr int NPAIR = (LCAP + 1) * (LCAP + 2) / 2;
    const float* base = A3f + (size_t)b.i * (8 * NPAIR);
    const double P0 = b.x1 * __ldg(p0tab + b.i) + b.x2 * __ldg(p0tab + b.i + 1);
    const double wlo = b.x1 * P0, whi = b.x2 * P0, fx2 = b.fac * b.x2;

    const double x = costh;
    const double somx2 = sqrt(BFE_MUL(BFE_SUB(1.0, x), BFE_ADD(1.0, x)));
    double xd = x;
    if (1.0 - fabs(xd) < 1.0e-8) xd = (xd > 0.0) ? (1.0 - 1.0e-8) : -(1.0 - 1.0e-8);
    const double dsom = BFE_DIV(1.0, BFE_SUB(BFE_MUL(xd, xd), 1.0));

    SlField f;
    f.pot0 = 0.0; f.pot1 = 0.0; f.potr = 0.0; f.pott = 0.0; f.potp = 0.0;
    double pmm = 1.0, fact = 1.0;
    double cm = 1.0, sm = 0.0;
    int offm = 0;
#pragma unroll
    for (int m = 0; m <= LCAP; ++m) {
        const int nl = LCAP - m + 1;
        if (m > 0) {
            pmm = BFE_MUL(pmm, BFE_MUL(-fact, somx2));
            fact += 2.0;
            double cn = cm * c1 - sm * s1, sn = sm * c1 + cm * s1;
            cm = cn; sm = sn;
        }
        double pl2 = 0.0, pl1 = pmm;
        double cl = cm, sl = sm;
#pragma unroll
        for (int l = m; l <= LCAP; ++l) {
            double P, dP;
            if (l == m) {
                P = pmm;
                dP = (l == 0) ? 0.0 : BFE_MUL(BFE_MUL(BFE_MUL(dsom, xd), (double)l), P);
            } else {
                if (l == m + 1) P = BFE_MUL(BFE_MUL(x, 2.0 * m + 1.0), pl1);
                else P = bfe_div_int(BFE_SUB(BFE_MUL(BFE_MUL(x, (double)(2 * l - 1)), pl1),
                                             BFE_MUL((double)(l + m - 1), pl2)), l - m);
                dP = BFE_MUL(dsom, BFE_SUB(BFE_MUL(BFE_MUL(xd, (double)l), P), BFE_MUL((double)(l + m), pl1)));
                pl2 = pl1; pl1 = P;
                double cn = cl * c1 - sl * s1, sn = sl * c1 + cl * s1;
                cl = cn; sl = sn;
            }
            float e[8];
            bfe_ldg256f(base + 8 * (offm + l - m), e);
            const double fl = bfe_fac_at(fac, l * (LCAP + 1) + m);
            const double spc = wlo * (double)e[0] + whi * (double)e[2];
            const double sdc = fx2 * (double)e[6] + b.fac * (double)e[4];
            if (m == 0) {
                if (l == 0) {
                    f.pot0 = fl * spc;
                    f.potr += fl * sdc;
                } else {
                    f.pot1 += fl * P * spc;
                    f.potr += fl * P * sdc;
                    f.pott += fl * dP * spc;
                }
            } else {
                const double sps = wlo * (double)e[1] + whi * (double)e[3];
                const double sds = fx2 * (double)e[7] + b.fac * (double)e[5];
                const double ct = trig_index_l ? cl : cm;
                const double stt = trig_index_l ? sl : sm;
                const double Ap = spc * ct + sps * stt;
                const double Ad = sdc * ct + sds * stt;
                const double Bp = sps * ct - spc * stt;
                const double fP = fl * P;
                f.pot1 += fP * Ap;
                f.potr += fP * Ad;
                f.pott += fl * dP * Ap;
                f.potp += fP * (double)m * Bp;
            }
        }
        offm += nl;
    }
    return f;
}

// potential.py:475-497 (Cartesian) / 425-440 (cylindrical): disc + halo field components -> the 8 outputs.
// The ten quotients by r2, r2^2, r3, r3^3 are formed from two reciprocals (a few ulp from the reference's
// separate divisions: these are plain products, nothing downstream amplifies them).
// The disc half and the halo half of the eight outputs depend on one expansion each, so the key-ordered point path can
// evaluate them in two kernels (bfe_orbit_sort.cu); bfe_cart_combine is the two halves together -- same expressions.
struct CartHalf { double fx, fy, fz, p; };

template <bool CYL>
__device__ __forceinline__ CartHalf bfe_cart_disc(const EofField& d, double x, double y, double r2, double r3, double xi0) {
    double diskfr = d.fr, diskfp = d.fp, diskfz = d.fz, diskp = d.p + d.p0;
    if (r3 < xi0) diskfp = 0.0;                              // 483-485 (min(xi) = xi[0])
    CartHalf o;
    if (CYL) {
        o.fx = diskfr;
        o.fy = diskfp;
        o.fz = diskfz;
        o.p = -1.0 * diskp;                                  // 440
        return o;
    }
    const double ir2 = 1.0 / r2;
    const double ir2sq = ir2 * ir2;
    o.fx = diskfr * (x * ir2) - diskfp * (y * ir2sq);
    o.fy = diskfr * (y * ir2) + diskfp * (x * ir2sq);
    o.fz = diskfz;
    o.p = diskp;
    return o;
}

template <bool CYL>
__device__ __forceinline__ CartHalf bfe_cart_halo(const SlField& h, double x, double y, double z, double r2, double r3,
                                                  double xi0) {
    double halofr = h.potr, haloft = h.pott, halofp = h.potp;
    if (r3 < xi0) halofp = 0.0;                              // 483-485
    CartHalf o;
    const double ir3 = 1.0 / r3;
    if (CYL) {
        o.fx = -1.0 * (r2 * halofr + z * haloft) * ir3;      // 433
        o.fy = -1.0 * halofp;
        o.fz = -1.0 * (z * halofr - r2 * haloft) * ir3;      // 435
        o.p = h.pot1 + h.pot0;
        return o;
    }
    const double ir2 = 1.0 / r2;
    const double ir2sq = ir2 * ir2, ir3cu = ir3 * ir3 * ir3;
    o.fx = -1.0 * (halofr * (x * ir3) - haloft * (x * z * ir3cu)) + halofp * (y * ir2sq);
    o.fy = -1.0 * (halofr * (y * ir3) - haloft * (y * z * ir3cu)) - halofp * (x * ir2sq);
    o.fz = -1.0 * (halofr * (z * ir3) + haloft * (r2 * r2 * ir3cu));
    o.p = h.pot1 + h.pot0;
    return o;
}

template <bool CYL>
__device__ __forceinline__ CartForce bfe_cart_combine(const EofField& d, const SlField& h, double x, double y, double z,
                                                      double r2, double r3, double xi0) {
    const CartHalf a = bfe_cart_disc<CYL>(d, x, y, r2, r3, xi0);
    const CartHalf b = bfe_cart_halo<CYL>(h, x, y, z, r2, r3, xi0);
    CartForce o;
    o.fxd = a.fx; o.fxh = b.fx; o.fyd = a.fy; o.fyh = b.fy; o.fzd = a.fz; o.fzh = b.fz; o.pd = a.p; o.ph = b.p;
    return o;
}

// CYL = true gives Fields.return_forces_cyl (potential.py:389-440) in the same 8 slots:
// diskfr, frhalo, diskfp, -halofp, diskfz, fzhalo, -diskp, halop+halop0.
template <int MCAP, int LCAP, bool CYL = false>
__device__ __forceinline__ CartForce bfe_field_cart(const EofGeom& ge, const double* __restrict__ G, int gstride,
                                                    const SlGeom& gs, const double2* __restrict__ A, int kpad,
                                                    const double* __restrict__ xi, const double* __restrict__ p0tab,
                                                    const double* __restrict__ fac,
                                                    double x, double y, double z, double crot, double srot) {
    const double eps = CYL ? 1.e-10 : 1.e-15;                               // 399-400 / 455-456
    double r2 = sqrt(BFE_ADD(BFE_MUL(x, x), BFE_MUL(y, y))) + eps;
    double r3 = sqrt(BFE_ADD(BFE_MUL(r2, r2), BFE_MUL(z, z))) + eps;
    double costh = BFE_DIV(z, r3);                                          // 457
    double c1, s1;
    bfe_cossin_phi(x, y, c1, s1);                           // 458
    // phi + rotpos
    double cr = c1 * crot - s1 * srot;
    double sr = s1 * crot + c1 * srot;
    EofBin eb = bfe_eof_bin(ge, r2, z);
    EofField d = bfe_eof_eval<MCAP>(ge, G, gstride, eb, cr, sr);
    SlBin sb = bfe_sl_bin(gs, xi, r3);
    SlField h = bfe_sl_eval<LCAP>(gs, A, kpad, p0tab, fac, sb, costh, cr, sr, true);
    return bfe_cart_combine<CYL>(d, h, x, y, z, r2, r3, gs.xi0);
}

// The same with the block evaluations above (G4 / A3, 256-bit loads); valid for g.lmax == LCAP.
// Split in two so that a kernel can stage the blocks a tile of points needs between the halves: the prologue finds
// the bins, the epilogue evaluates from two block base pointers.
struct FieldPt {
    double x, y, z, r2, r3, costh, cr, sr;
    EofBin eb;
    SlBin sb;
};

template <bool CYL>
__device__ __forceinline__ FieldPt bfe_field_prologue(const EofGeom& ge, const SlGeom& gs, const double* __restrict__ xi,
                                                      double x, double y, double z, double crot, double srot) {
    const double eps = CYL ? 1.e-10 : 1.e-15;
    FieldPt p;
    p.x = x; p.y = y; p.z = z;
    p.r2 = sqrt(BFE_ADD(BFE_MUL(x, x), BFE_MUL(y, y))) + eps;
    p.r3 = sqrt(BFE_ADD(BFE_MUL(p.r2, p.r2), BFE_MUL(z, z))) + eps;
    p.costh = BFE_DIV(z, p.r3);
    double c1, s1;
    bfe_cossin_phi(x, y, c1, s1);
    p.cr = c1 * crot - s1 * srot;
    p.sr = s1 * crot + c1 * srot;
    p.eb = bfe_eof_bin(ge, p.r2, z);
    p.sb = bfe_sl_bin(gs, xi, p.r3);
    return p;
}

// FP64 tables: baseE = block of cell p.eb.cell, baseS = block of radial index max(p.sb.i, 1)
template <int MCAP, int LCAP, bool CYL, typename FacT, class LD>
__device__ __forceinline__ CartForce bfe_field_epilogue(const EofGeom& ge, const SlGeom& gs, const double2* baseE,
                                                        const double2* baseS, const double* __restrict__ p0tab,
                                                        const FacT& fac, const FieldPt& p) {
    const EofField d = bfe_eof_eval_base<MCAP, LD>(ge, baseE, p.eb, p.cr, p.sr);
    const SlField h = bfe_sl_eval_base<LCAP, FacT, LD>(gs, baseS, p0tab, fac, p.sb, p.costh, p.cr, p.sr, true);
    return bfe_cart_combine<CYL>(d, h, p.x, p.y, p.z, p.r2, p.r3, gs.xi0);
}

template <int MCAP, int LCAP, bool CYL = false, bool F32 = false, typename FacT = const double*>
__device__ __forceinline__ CartForce bfe_field_cart_blk(const EofGeom& ge, const void* __restrict__ G4,
                                                        const SlGeom& gs, const void* __restrict__ A3,
                                                        const double* __restrict__ xi, const double* __restrict__ p0tab,
                                                        const FacT& fac,
                                                        double x, double y, double z, double crot, double srot) {
    const FieldPt p = bfe_field_prologue<CYL>(ge, gs, xi, x, y, z, crot, srot);
    if constexpr (F32) {
        const EofField d = bfe_eof_eval_blk32<MCAP>(ge, static_cast<const float*>(G4), p.eb, p.cr, p.sr);
        const SlField h = bfe_sl_eval_blk32<LCAP>(gs, static_cast<const float*>(A3), p0tab, fac, p.sb, p.costh, p.cr, p.sr, true);
        return bfe_cart_combine<CYL>(d, h, x, y, z, p.r2, p.r3, gs.xi0);
    } else {
        // FP64: EOF block of the cell (G4), SL polynomial block of the interval (A4)
        const EofField d = bfe_eof_eval_base<MCAP, LdGlobal256>(
            ge, static_cast<const double2*>(G4) + (size_t)p.eb.cell * (size_t)(12 * (ge.mmax + 1)), p.eb, p.cr, p.sr);
        const SlField h = bfe_sl_eval_poly<LCAP>(gs, A3, p0tab, p.sb, p.costh, p.cr, p.sr, true);
        return bfe_cart_combine<CYL>(d, h, x, y, z, p.r2, p.r3, gs.xi0);
    }
}

template <int MCAP, int LCAP, bool CYL>
__device__ __forceinline__ CartForce bfe_field_cart_staged(const EofGeom& ge, const double2* __restrict__ G4,
                                                           const SlGeom& gs, const double2* __restrict__ A3,
                                                           const double* __restrict__ xi, const double* __restrict__ p0tab,
                                                           const double* __restrict__ fac, double x, double y, double z,
                                                           double crot, double srot, double2* __restrict__ st, int lane) {
    const double eps = CYL ? 1.e-10 : 1.e-15;
    double r2 = sqrt(BFE_ADD(BFE_MUL(x, x), BFE_MUL(y, y))) + eps;
    double r3 = sqrt(BFE_ADD(BFE_MUL(r2, r2), BFE_MUL(z, z))) + eps;
    double costh = BFE_DIV(z, r3);
    double c1, s1;
    bfe_cossin_phi(x, y, c1, s1);
    double cr = c1 * crot - s1 * srot;
    double sr = s1 * crot + c1 * srot;
    EofBin eb = bfe_eof_bin(ge, r2, z);
    EofField d = bfe_eof_eval_staged<MCAP>(ge, G4, eb, cr, sr, st, lane);
    SlBin sb = bfe_sl_bin(gs, xi, r3);
    SlField h = bfe_sl_eval_staged<LCAP>(gs, A3, p0tab, fac, sb, costh, cr, sr, true, st, lane);
    return bfe_cart_combine<CYL>(d, h, x, y, z, r2, r3, gs.xi0);
}
