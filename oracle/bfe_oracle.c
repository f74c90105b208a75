/*
 * bfe_oracle.c -- plain-C restatement (FP64, OpenMP over particles) of the accumulate / force
 * passes of exptool's BFE hot path, in the reference's DIRECT formulation (no contraction over n,
 * no sorting): every particle interpolates every (m,n) table, exactly as eof.py / spheresl.py do.
 *
 * TEST INFRASTRUCTURE ONLY.  Built into oracle/_build/libbfe_oracle.so by oracle/build_oracle.py;
 * loaded only by tests/ and by bench.py's cpu_baseline / --impl reference legs (through
 * oracle/oracle_c.py).  The product never links or loads it.
 *
 * Pinned through the NumPy oracle: tests/test_oracle_c.py asserts this file agrees with
 * oracle/oracle_np.py (itself pinned to the reference's golden vectors) to <= 1e-12.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FOURPI_NEG (-12.566370614359172953850573533118)

/* compatibility.py:16-47 */
static double r_to_xi(double r, int cmap, double scale) {
    double out;
    if (cmap == 1) out = (r / scale - 1.0) / (r / scale + 1.0);
    else if (cmap == 2) out = log(r);
    else out = r;
    return (r < 0.0) ? 0.0 : out;
}
/* compatibility.py:73-80 */
static double d_xi_to_r(double xi, int cmap, double scale) {
    if (cmap == 1) return 0.5 * (1.0 - xi) * (1.0 - xi) / scale;
    if (cmap == 2) return exp(-xi);
    return 1.0;
}
/* compatibility.py:91 */
static double z_to_y(double z, double hscale) { return (z / (fabs(z) + 1.e-8)) * asinh(fabs(z / hscale)); }

typedef struct {
    int mmax, norder, numx, numy, cmap;
    double xmin, dx, ymin, dy, ascale, hscale;
} eof_geom;

/* eof.py:394-422 + 443-451: bins (truncation, lower clamp, upper edge extrapolates) and weights */
static void eof_bins(const eof_geom* g, double r, double z, int* ix, int* iy, double c[4]) {
    double X = (r_to_xi(r, g->cmap, g->ascale) - g->xmin) / g->dx;
    double Y = (z_to_y(z, g->hscale) - g->ymin) / g->dy;
    int i = (int)X, j = (int)Y;
    if (i < 0) i = 0;
    if (X < 0) X = 0;
    if (i >= g->numx) i = g->numx - 1;
    if (j < 0) j = 0;
    if (Y < 0) Y = 0;
    if (j >= g->numy) j = g->numy - 1;
    double dx0 = i + 1.0 - X, dy0 = j + 1.0 - Y, dx1 = X - i, dy1 = Y - j;
    c[0] = dx0 * dy0; c[1] = dx1 * dy0; c[2] = dx0 * dy1; c[3] = dx1 * dy1;
    *ix = i; *iy = j;
}

static inline double interp(const double* T, int ny1, int ix, int iy, const double c[4]) {
    const double* p = T + (size_t)ix * ny1 + iy;
    return p[0] * c[0] + p[ny1] * c[1] + p[1] * c[2] + p[ny1 + 1] * c[3];
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline must still use every host core */
void bfe_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int bfe_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* eof.accumulate, eof.py:526-551.  potC/potS: [m][n][numx+1][numy+1]; out: (mmax+1)*norder each */
void bfe_oracle_eof_accumulate(const eof_geom* g, const double* potC, const double* potS, long n, const double* x,
                               const double* y, const double* z, const double* mass, double* cos_out,
                               double* sin_out) {
    const int M = g->mmax + 1, N = g->norder, ny1 = g->numy + 1;
    const size_t plane = (size_t)(g->numx + 1) * ny1;
    const int nc = M * N;
    memset(cos_out, 0, sizeof(double) * nc);
    memset(sin_out, 0, sizeof(double) * nc);
#pragma omp parallel
    {
        double* ac = (double*)calloc(2 * nc, sizeof(double));
        double* as = ac + nc;
#pragma omp for schedule(static)
        for (long p = 0; p < n; ++p) {
            double r = sqrt(x[p] * x[p] + y[p] * y[p] + 1.e-10);   /* 531 */
            double phi = atan2(y[p], x[p]);                        /* 532 */
            int ix, iy;
            double c[4];
            eof_bins(g, r, z[p], &ix, &iy, c);
            for (int m = 0; m < M; ++m) {
                double cm = cos(phi * m), sm = sin(phi * m);       /* 541-542 */
                for (int k = 0; k < N; ++k) {
                    size_t off = ((size_t)m * N + k) * plane;
                    double vc = interp(potC + off, ny1, ix, iy, c) * mass[p];
                    double vs = interp(potS + off, ny1, ix, iy, c) * mass[p];
                    ac[m * N + k] += FOURPI_NEG * cm * vc;         /* 550 */
                    as[m * N + k] += FOURPI_NEG * sm * vs;
                }
            }
        }
#pragma omp critical
        for (int k = 0; k < nc; ++k) { cos_out[k] += ac[k]; sin_out[k] += as[k]; }
        free(ac);
    }
}

/* eof.accumulated_eval_particles, eof.py:1068-1138.  tabs: potC,rfC,zfC,potS,rfS,zfS */
void bfe_oracle_eof_force(const eof_geom* g, const double* const tabs[6], const double* cosc, const double* sinc,
                          int m1, int m2, long n, const double* x, const double* y, const double* z,
                          double* p0, double* pp, double* fr, double* fp, double* fz, double* R) {
    const int M = g->mmax + 1, N = g->norder, ny1 = g->numy + 1;
    const size_t plane = (size_t)(g->numx + 1) * ny1;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < n; ++p) {
        double r = sqrt(x[p] * x[p] + y[p] * y[p] + 1.e-10);       /* 1070 */
        double phi = atan2(y[p], x[p]);
        int ix, iy;
        double c[4];
        eof_bins(g, r, z[p], &ix, &iy, c);
        double sp0 = 0, sp = 0, sfr = 0, sfp = 0, sfz = 0;
        for (int m = 0; m < M; ++m) {
            if (m > m2 || m < m1) continue;                        /* 1094 */
            double cm = cos(phi * m), sm = sin(phi * m);
            double vp = 0, vr = 0, vz = 0, wp = 0, wr = 0, wz = 0;
            for (int k = 0; k < N; ++k) {
                size_t off = ((size_t)m * N + k) * plane;
                double a = cosc[m * N + k];
                vp += a * interp(tabs[0] + off, ny1, ix, iy, c);
                vr += a * interp(tabs[1] + off, ny1, ix, iy, c);
                vz += a * interp(tabs[2] + off, ny1, ix, iy, c);
                if (m > 0) {
                    double b = sinc[m * N + k];
                    wp += b * interp(tabs[3] + off, ny1, ix, iy, c);
                    wr += b * interp(tabs[4] + off, ny1, ix, iy, c);
                    wz += b * interp(tabs[5] + off, ny1, ix, iy, c);
                }
            }
            sfr += cm * vr + sm * wr;
            sfz += cm * vz + sm * wz;
            sfp += m * (sm * vp - cm * wp);
            if (m == 0) sp0 = cm * vp;                             /* 1129-1134 */
            else sp += cm * vp + sm * wp;
        }
        p0[p] = sp0; pp[p] = sp; fr[p] = sfr; fp[p] = sfp; fz[p] = sfz; R[p] = r;
    }
}

typedef struct {
    int lmax, nmax, numr, cmap;
    double scale;
} sl_geom;

/* spheresl.legendre_R, spheresl.py:664-700 */
static void legendre(int lmax, double x, double* P /* (lmax+1)^2 */) {
    const int L = lmax + 1;
    memset(P, 0, sizeof(double) * L * L);
    P[0] = 1.0;
    double pll = 1.0;
    if (lmax > 0) {
        double somx2 = sqrt((1.0 - x) * (1.0 + x)), fact = 1.0;
        for (int m = 1; m <= lmax; ++m) { pll *= -fact * somx2; P[m * L + m] = pll; fact += 2.0; }
    }
    for (int m = 0; m < lmax; ++m) {
        double pl2 = P[m * L + m];
        double pl1 = x * (2. * m + 1) * pl2;
        P[(m + 1) * L + m] = pl1;
        for (int l = m + 2; l <= lmax; ++l) {
            double v = (x * (2 * l - 1) * pl1 - (l + m - 1) * pl2) / (l - m);
            P[l * L + m] = v;
            pl2 = pl1; pl1 = v;
        }
    }
}

/* spheresl.compute_coefficients_solitary, spheresl.py:586-654.  fac: factorial_return (lmax+1)^2 */
void bfe_oracle_sl_accumulate(const sl_geom* g, const double* ev, const double* ef, const double* xi,
                              const double* p0, const double* fac, int no_odd, long n, const double* x,
                              const double* y, const double* z, const double* mass, double* expcoef) {
    const int L = g->lmax + 1, N = g->nmax, nr = g->numr;
    const int nc = L * L * N;
    const double dxi = xi[1] - xi[0];
    memset(expcoef, 0, sizeof(double) * nc);
#pragma omp parallel
    {
        double* acc = (double*)calloc(nc, sizeof(double));
        double* P = (double*)malloc(sizeof(double) * L * L);
        double* potd = (double*)malloc(sizeof(double) * L * N);
#pragma omp for schedule(static)
        for (long p = 0; p < n; ++p) {
            double r2 = x[p] * x[p] + y[p] * y[p] + z[p] * z[p];
            double r = fmax(sqrt(r2), 1.0e-10);                    /* 611 */
            double costh = z[p] / r, phi = atan2(y[p], x[p]);
            legendre(g->lmax, costh, P);
            double xx = r_to_xi(r, g->cmap, g->scale);             /* 309-328 */
            if (g->cmap == 1) { if (xx < -1.0) xx = -1.0; if (xx >= 1.0) xx = 1.0 - 1.0e-08; }
            int i = (int)floor((xx - xi[0]) / dxi);
            if (i < 0) i = 0;
            if (i > nr - 2) i = nr - 2;
            double x1 = (xi[i + 1] - xx) / dxi, x2 = (xx - xi[i]) / dxi;
            double P0 = x1 * p0[i] + x2 * p0[i + 1];
            for (int l = 0; l < L; ++l)
                for (int k = 0; k < N; ++k) {
                    const double* e = ef + ((size_t)l * N + k) * nr;
                    potd[l * N + k] = (x1 * e[i] + x2 * e[i + 1]) / sqrt(ev[l * N + k]) * P0;   /* 332 */
                }
            int loffset = 0;
            for (int l = 0; l < L; ++l) {
                if ((l % 2) != 0 && no_odd) { loffset += 2 * l + 1; continue; }
                int moffset = 0;
                for (int m = 0; m <= l; ++m) {
                    double f = fac[l * L + m] * P[l * L + m];
                    if (m == 0) {
                        for (int k = 0; k < N; ++k) acc[(loffset + moffset) * N + k] += potd[l * N + k] * f * FOURPI_NEG * mass[p];
                        moffset += 1;
                    } else {
                        double cm = cos(phi * m), sm = sin(phi * m);
                        for (int k = 0; k < N; ++k) {
                            double f4 = potd[l * N + k] * f * FOURPI_NEG;
                            acc[(loffset + moffset) * N + k] += cm * f4 * mass[p];
                            acc[(loffset + moffset + 1) * N + k] += sm * f4 * mass[p];
                        }
                        moffset += 2;
                    }
                }
                loffset += 2 * l + 1;
            }
        }
#pragma omp critical
        for (int k = 0; k < nc; ++k) expcoef[k] += acc[k];
        free(acc); free(P); free(potd);
    }
    (void)d_xi_to_r;
}
