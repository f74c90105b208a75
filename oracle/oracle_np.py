"""
oracle_np.py -- CPU restatement (NumPy, FP64) of exptool's BFE hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product
(exptool_b200/) never imports it and has no CPU fallback.

Parity status: the reference ships no golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF:
tests/golden/make_golden.py runs the unmodified reference functions on the
synthetic table files and freezes the results in tests/golden/*.npz;
tests/test_oracle_golden.py asserts this file reproduces them (<= 1e-13
relative), and tests/test_oracle_vs_reference.py repeats the comparison live
whenever /root/reference is present.

Every function cites the reference lines it restates.  The restatement is
vectorised over particles (the reference loops in Python for most of them), which
changes only FP64 summation order.
"""
import numpy as np
from scipy.special import gammaln

FOURPI_NEG = -4.0 * np.pi


# ---------------------------------------------------------------------------
# coordinate maps -- exptool/basis/compatibility.py:16-99
# ---------------------------------------------------------------------------
def r_to_xi(r, cmap, scale):
    """compatibility.py:16-47 (negatives -> 0; the cmap=0 in-place mutation is not replicated)."""
    r = np.asarray(r, dtype=np.float64)
    if cmap == 1:
        out = (r / scale - 1.0) / (r / scale + 1.0)
    elif cmap == 2:
        with np.errstate(invalid='ignore', divide='ignore'):
            out = np.log(r)
    else:
        out = r.copy()
    return np.where(r < 0.0, 0.0, out)


def xi_to_r(xi, cmap, scale):
    """compatibility.py:54-62"""
    if cmap == 1:
        return (1.0 + xi) / (1.0 - xi) * scale
    if cmap == 2:
        return np.exp(xi)
    return xi


def d_xi_to_r(xi, cmap, scale):
    """compatibility.py:65-80"""
    if cmap == 1:
        return 0.5 * (1.0 - xi) * (1.0 - xi) / scale
    if cmap == 2:
        return np.exp(-xi)
    return np.ones_like(xi)


def z_to_y(z, hscale):
    """compatibility.py:83-91 (epsilon 1e-8, the live Python value)."""
    z = np.asarray(z, dtype=np.float64)
    return (z / (np.abs(z) + 1.0e-8)) * np.arcsinh(np.abs(z / hscale))


# ---------------------------------------------------------------------------
# EOF geometry -- eof.py:316-347, 354-427
# ---------------------------------------------------------------------------
def eof_set_table_params(RMAX=20.0, RMIN=0.001, ASCALE=0.01, HSCALE=0.001, NUMX=128, NUMY=64, CMAP=0):
    """eof.py:316-347"""
    Rtable = np.sqrt(0.5) * RMAX
    XMIN = float(r_to_xi(RMIN * ASCALE, CMAP, ASCALE))
    XMAX = float(r_to_xi(Rtable * ASCALE, CMAP, ASCALE))
    dX = (XMAX - XMIN) / NUMX
    YMIN = float(z_to_y(-Rtable * ASCALE, HSCALE))
    YMAX = float(z_to_y(Rtable * ASCALE, HSCALE))
    dY = (YMAX - YMIN) / NUMY
    return XMIN, XMAX, dX, YMIN, YMAX, dY


def eof_return_bins(r, z, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP):
    """
    eof.py:394-422.  Truncating int cast; lower edge clamps X and ix; upper edge
    clamps ix only (the second mask at 414-415 / 421-422 is empty after the first
    assignment) so the interpolation extrapolates beyond the table.
    """
    X = (r_to_xi(r, CMAP, ASCALE) - rmin) / dR
    Y = (z_to_y(z, HSCALE) - zmin) / dZ
    ix = X.astype(np.int64)
    iy = Y.astype(np.int64)
    ix[ix < 0] = 0
    X[X < 0] = 0.0
    ix[ix >= numx] = numx - 1
    iy[iy < 0] = 0
    Y[Y < 0] = 0.0
    iy[iy >= numy] = numy - 1
    return X, Y, ix, iy


def _bilinear_weights(X, Y, ix, iy):
    """eof.py:443-451"""
    delx0 = ix + 1.0 - X
    dely0 = iy + 1.0 - Y
    delx1 = X - ix
    dely1 = Y - iy
    return delx0 * dely0, delx1 * dely0, delx0 * dely1, delx1 * dely1


def _interp(T, ix, iy, c00, c10, c01, c11):
    """T[..., ix, iy] bilinear; T shape (..., numx+1, numy+1) -> (..., N)."""
    return (T[..., ix, iy] * c00 + T[..., ix + 1, iy] * c10 +
            T[..., ix, iy + 1] * c01 + T[..., ix + 1, iy + 1] * c11)


# ---------------------------------------------------------------------------
# EOF accumulate -- eof.py:492-551 (+430-457)
# ---------------------------------------------------------------------------
def eof_accumulate(x, y, z, mass, potC, potS, MMAX, NMAX, XMIN, dX, YMIN, dY,
                   NUMX, NUMY, ASCALE, HSCALE, CMAP, chunk=16384):
    """
    cos[m,n] = sum_p -4pi cos(m phi_p) (m_p interp_p(potC[m,n])), sin likewise
    (eof.py:526-551).  r = sqrt(x^2+y^2+1e-10) (531); no_odd has no effect in the
    reference (mask computed, never applied) so it is not a parameter here.
    """
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    z = np.asarray(z, np.float64); mass = np.asarray(mass, np.float64)
    acc_c = np.zeros((MMAX + 1, NMAX))
    acc_s = np.zeros((MMAX + 1, NMAX))
    morder = np.arange(MMAX + 1, dtype=np.float64)[:, None]
    pc = potC[:MMAX + 1, :NMAX]
    ps = potS[:MMAX + 1, :NMAX]
    for lo in range(0, x.size, chunk):
        sl = slice(lo, lo + chunk)
        r = (x[sl] ** 2. + y[sl] ** 2. + 1.e-10) ** 0.5
        phi = np.arctan2(y[sl], x[sl])
        X, Y, ix, iy = eof_return_bins(r, z[sl], XMIN, dX, YMIN, dY, NUMX, NUMY, ASCALE, HSCALE, CMAP)
        c = _bilinear_weights(X, Y, ix, iy)
        vc = _interp(pc, ix, iy, *c) * mass[sl]
        vs = _interp(ps, ix, iy, *c) * mass[sl]
        mcos = np.cos(phi[None, :] * morder)[:, None, :]
        msin = np.sin(phi[None, :] * morder)[:, None, :]
        acc_c += np.sum(FOURPI_NEG * mcos * vc, axis=2)
        acc_s += np.sum(FOURPI_NEG * msin * vs, axis=2)
    return acc_c, acc_s


def partition_like_reference(n, divisions):
    """eof.py:1336-1354 / spheresl.py:383-403: chunk 0 takes the remainder."""
    avg = int(np.floor(n / divisions))
    first = n - avg * (divisions - 1)
    bounds = [0, first]
    for _ in range(1, divisions):
        bounds.append(bounds[-1] + avg)
    return bounds


# ---------------------------------------------------------------------------
# EOF field evaluation
# ---------------------------------------------------------------------------
def eof_force_particles(x, y, z, accum_cos, accum_sin, potC, rforceC, zforceC,
                        potS, rforceS, zforceS, rmin, dR, zmin, dZ, numx, numy,
                        MMAX, NMAX, ASCALE, HSCALE, CMAP, m1=0, m2=1000, chunk=16384, densC=None, densS=None):
    """
    eof.accumulated_eval_particles (eof.py:989-1144), vectorised over particles.
    Returns p0, p, fr, fp, fz, R with p excluding m=0 (1129-1134), fr/fz including
    it, R = sqrt(x^2+y^2+1e-10) (1070), window m1 <= m <= m2 (1094).
    With densC/densS (density=True, eof.py:1106,1122,1136-1138) returns p0, p, d0, d, fr, fp, fz, R.
    """
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64); z = np.asarray(z, np.float64)
    n = x.size
    p0 = np.zeros(n); p = np.zeros(n); fr = np.zeros(n); fp = np.zeros(n); fz = np.zeros(n)
    d0 = np.zeros(n); d = np.zeros(n)
    R = (x * x + y * y + 1.e-10) ** 0.5
    PHI = np.arctan2(y, x)
    for lo in range(0, n, chunk):
        sl = slice(lo, lo + chunk)
        X, Y, ix, iy = eof_return_bins(R[sl], z[sl], rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP)
        c = _bilinear_weights(X, Y, ix, iy)
        phi = PHI[sl]
        for mm in range(0, MMAX + 1):
            if (mm > m2) or (mm < m1):
                continue
            ccos = np.cos(phi * mm)
            ssin = np.sin(phi * mm)
            ac = accum_cos[mm, :NMAX, None]
            vp = np.sum(ac * _interp(potC[mm, :NMAX], ix, iy, *c), axis=0)
            vr = np.sum(ac * _interp(rforceC[mm, :NMAX], ix, iy, *c), axis=0)
            vz = np.sum(ac * _interp(zforceC[mm, :NMAX], ix, iy, *c), axis=0)
            pm = ccos * vp
            dm = ccos * np.sum(ac * _interp(densC[mm, :NMAX], ix, iy, *c), axis=0) if densC is not None else 0.0
            fr[sl] += ccos * vr
            fz[sl] += ccos * vz
            fp[sl] += ssin * mm * vp
            if mm > 0:
                asn = accum_sin[mm, :NMAX, None]
                wp = np.sum(asn * _interp(potS[mm, :NMAX], ix, iy, *c), axis=0)
                wr = np.sum(asn * _interp(rforceS[mm, :NMAX], ix, iy, *c), axis=0)
                wz = np.sum(asn * _interp(zforceS[mm, :NMAX], ix, iy, *c), axis=0)
                pm = pm + ssin * wp
                if densS is not None:
                    dm = dm + ssin * np.sum(asn * _interp(densS[mm, :NMAX], ix, iy, *c), axis=0)
                fr[sl] += ssin * wr
                fz[sl] += ssin * wz
                fp[sl] += -ccos * mm * wp
                p[sl] += pm
                d[sl] += dm
            else:
                p0[sl] = pm
                d0[sl] = dm
    if densC is not None:
        return p0, p, d0, d, fr, fp, fz, R
    return p0, p, fr, fp, fz, R


def eof_force_eval(r, z, phi, accum_cos, accum_sin, potC, rforceC, zforceC,
                   potS, rforceS, zforceS, rmin, dR, zmin, dZ, numx, numy,
                   MMAX, NMAX, ASCALE, HSCALE, CMAP, no_odd=False, perturb=False):
    """
    eof.force_eval (eof.py:756-870), vectorised over points (r,z,phi arrays).
    MMAX/NMAX slice coefficients and tables (784-793); no_odd mask is
    |cos(pi m/2)| (832-833).  Returns (fr+fr0, fp, fz+fz0, p+p0, p0), or the
    7-tuple if perturb.
    """
    r = np.atleast_1d(np.asarray(r, np.float64)); z = np.atleast_1d(np.asarray(z, np.float64))
    phi = np.atleast_1d(np.asarray(phi, np.float64))
    X, Y, ix, iy = eof_return_bins(r.copy(), z, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP)
    c = _bilinear_weights(X, Y, ix, iy)
    n = r.size
    p = np.zeros(n); fr = np.zeros(n); fz = np.zeros(n); fp = np.zeros(n)
    a0 = accum_cos[0, :NMAX, None]
    p0 = np.sum(a0 * _interp(potC[0, :NMAX], ix, iy, *c), axis=0)
    fr0 = np.sum(a0 * _interp(rforceC[0, :NMAX], ix, iy, *c), axis=0)
    fz0 = np.sum(a0 * _interp(zforceC[0, :NMAX], ix, iy, *c), axis=0)
    for mm in range(1, MMAX + 1):
        mask = abs(np.cos((np.pi / 2.) * mm)) if no_odd else 1.0
        ccos = np.cos(phi * mm); ssin = np.sin(phi * mm)
        ac = accum_cos[mm, :NMAX, None]; asn = accum_sin[mm, :NMAX, None]
        vp = np.sum(ac * _interp(potC[mm, :NMAX], ix, iy, *c), axis=0)
        vr = np.sum(ac * _interp(rforceC[mm, :NMAX], ix, iy, *c), axis=0)
        vz = np.sum(ac * _interp(zforceC[mm, :NMAX], ix, iy, *c), axis=0)
        wp = np.sum(asn * _interp(potS[mm, :NMAX], ix, iy, *c), axis=0)
        wr = np.sum(asn * _interp(rforceS[mm, :NMAX], ix, iy, *c), axis=0)
        wz = np.sum(asn * _interp(zforceS[mm, :NMAX], ix, iy, *c), axis=0)
        p += mask * (ccos * vp + ssin * wp)
        fr += mask * (ccos * vr + ssin * wr)
        fz += mask * (ccos * vz + ssin * wz)
        fp += mask * mm * (ssin * vp - ccos * wp)
    if perturb:
        return fr, fp, fz, p, p0, fr0, fz0
    return fr + fr0, fp, fz + fz0, p + p0, p0


def eof_get_pot(r, z, cos_array, sin_array, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP, fac=1.0):
    """eof.get_pot (eof.py:430-457) -> Vc, Vs of shape (m, n, N)."""
    r = np.atleast_1d(np.asarray(r, np.float64)); z = np.atleast_1d(np.asarray(z, np.float64))
    X, Y, ix, iy = eof_return_bins(r.copy(), z, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP)
    c = _bilinear_weights(X, Y, ix, iy)
    return fac * _interp(cos_array, ix, iy, *c), fac * _interp(sin_array, ix, iy, *c)


def eof_accumulated_eval(r, z, phi, accum_cos, accum_sin, potC, rforceC, zforceC, densC, potS, rforceS, zforceS,
                         densS, rmin, dR, zmin, dZ, numx, numy, MMAX, NMAX, ASCALE, HSCALE, CMAP, no_odd=False):
    """
    eof.accumulated_eval (eof.py:874-929), vectorised over points: (p0, p, fr, fp, fz, d0, d) with p, d the
    TOTAL sums (m = 0 included, 925-927) and no_odd skipping odd m (896-897).
    """
    r = np.atleast_1d(np.asarray(r, np.float64)); z = np.atleast_1d(np.asarray(z, np.float64))
    phi = np.atleast_1d(np.asarray(phi, np.float64))
    X, Y, ix, iy = eof_return_bins(r.copy(), z, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP)
    c = _bilinear_weights(X, Y, ix, iy)
    n = r.size
    p = np.zeros(n); fr = np.zeros(n); fz = np.zeros(n); fp = np.zeros(n); d = np.zeros(n)
    p0 = np.zeros(n); d0 = np.zeros(n)
    for mm in range(0, MMAX + 1):
        if (mm % 2 != 0) and no_odd:
            continue
        ccos = np.cos(phi * mm); ssin = np.sin(phi * mm)
        ac = accum_cos[mm, :, None]; asn = accum_sin[mm, :, None]
        vp = np.sum(ac * _interp(potC[mm], ix, iy, *c), axis=0)
        p += ccos * vp
        fr += ccos * np.sum(ac * _interp(rforceC[mm], ix, iy, *c), axis=0)
        fz += ccos * np.sum(ac * _interp(zforceC[mm], ix, iy, *c), axis=0)
        d += ccos * np.sum(ac * _interp(densC[mm], ix, iy, *c), axis=0)
        fp += ssin * mm * vp
        if mm > 0:
            wp = np.sum(asn * _interp(potS[mm], ix, iy, *c), axis=0)
            p += ssin * wp
            fr += ssin * np.sum(asn * _interp(rforceS[mm], ix, iy, *c), axis=0)
            fz += ssin * np.sum(asn * _interp(zforceS[mm], ix, iy, *c), axis=0)
            d += ssin * np.sum(asn * _interp(densS[mm], ix, iy, *c), axis=0)
            fp += -ccos * mm * wp
        if mm == 0:
            p0 = p.copy(); d0 = d.copy()
    return p0, p, fr, fp, fz, d0, d


# ---------------------------------------------------------------------------
# SL tables -- halo_methods.py:178-220 (uses SciPy's own splrep/splev, as the
# reference does; see SURVEY.md section 8c "run SciPy itself")
# ---------------------------------------------------------------------------
def sl_init_table(R1, D1, P1, numr, rmin, rmax, cmap, scale):
    """halo_methods.init_table (178-220) from already-parsed model columns."""
    from scipy import interpolate
    if cmap == 1:
        xmin = (rmin / scale - 1.0) / (rmin / scale + 1.0)
        xmax = (rmax / scale - 1.0) / (rmax / scale + 1.0)
    elif cmap == 0:
        xmin, xmax = rmin, rmax
    else:
        raise ValueError('cmap=2 is broken in the reference (halo_methods.py:195-196)')
    dxi = (xmax - xmin) / (numr - 1)
    pfunc = interpolate.splrep(R1, P1, s=0)
    dfunc = interpolate.splrep(R1, 4. * np.pi * D1, s=0)
    xi = np.zeros(numr)
    for i in range(numr):
        xi[i] = xmin + dxi * i
    r = xi_to_r(xi, cmap, scale)
    p0 = interpolate.splev(r, pfunc, der=0)
    d0 = interpolate.splev(r, dfunc, der=0)
    return xi, r, p0, d0


def factorial_return(lmax):
    """spheresl.py:823-863"""
    f = np.zeros((lmax + 1, lmax + 1))
    for l in range(lmax + 1):
        for m in range(l + 1):
            f[l, m] = np.sqrt((0.5 * l + 0.25) / np.pi * np.exp(gammaln(1.0 + l - m) - gammaln(1.0 + l + m)))
            if m != 0:
                f[l, m] *= np.sqrt(2.)
    return f


def legendre_R(lmax, x):
    """spheresl.legendre_R (664-700) vectorised: returns p[l, m, N]; non-finite -> 0 (698)."""
    x = np.asarray(x, np.float64)
    p = np.zeros((lmax + 1, lmax + 1) + x.shape)
    p[0, 0] = 1.0
    pll = np.ones_like(x)
    with np.errstate(invalid='ignore', over='ignore'):
        if lmax > 0:
            somx2 = np.sqrt((1.0 - x) * (1.0 + x))
            fact = 1.0
            for m in range(1, lmax + 1):
                pll = pll * (-fact * somx2)
                p[m, m] = pll
                fact += 2.0
        for m in range(0, lmax):
            pl2 = p[m, m]
            pl1 = x * (2. * m + 1) * pl2
            p[m + 1, m] = pl1
            for l in range(m + 2, lmax + 1):
                pll = (x * (2 * l - 1) * pl1 - (l + m - 1) * pl2) / (l - m)
                p[l, m] = pll
                pl2 = pl1
                pl1 = pll
    p[~np.isfinite(p)] = 0.
    return p


def dlegendre_R(lmax, x):
    """spheresl.dlegendre_R (706-770) vectorised: (p, dp), each [l, m, N]."""
    x = np.asarray(x, np.float64)
    p = legendre_R(lmax, x)
    MINEPS = 1.e-8
    xx = x.copy()
    near = (1.0 - np.abs(xx)) < MINEPS
    xx = np.where(near, np.where(xx > 0, 1.0 - MINEPS, -(1.0 - MINEPS)), xx)
    dp = np.zeros_like(p)
    with np.errstate(invalid='ignore', divide='ignore', over='ignore'):
        somx2 = 1.0 / (xx * xx - 1.0)
        for l in range(1, lmax + 1):
            for m in range(0, l):
                dp[l, m] = somx2 * (xx * l * p[l, m] - (l + m) * p[l - 1, m])
            dp[l, l] = somx2 * xx * l * p[l, l]
    dp[~np.isfinite(dp)] = 0.
    return p, dp


def _sl_bins(r, xi, cmap, scale):
    """spheresl.py:123-140 / 309-328: xi clamp, floor bin, clamp to [0, numr-2], x1, x2."""
    numr = xi.shape[0]
    x = r_to_xi(r, cmap, scale)
    if cmap == 1:
        x = np.where(x < -1.0, -1.0, x)
        x = np.where(x >= 1.0, 1.0 - 1.0e-08, x)
    dxi = xi[1] - xi[0]
    indx = np.floor((x - np.min(xi)) / dxi).astype(np.int64)
    indx = np.clip(indx, 0, numr - 2)
    x1 = (xi[indx + 1] - x) / dxi
    x2 = (x - xi[indx]) / dxi
    return x, dxi, indx, x1, x2


def sl_pot_matrix(r, lmax, nmax, evtable, eftable, xi, p0, cmap, scale):
    """spheresl.get_halo_pot_matrix (301-335) vectorised -> potd[l, n, N]."""
    x, dxi, i, x1, x2 = _sl_bins(r, xi, cmap, scale)
    ef = eftable[:lmax + 1, :nmax]
    sq = np.sqrt(evtable[:lmax + 1, :nmax])[:, :, None]
    return (x1 * ef[:, :, i] + x2 * ef[:, :, i + 1]) / sq * (x1 * p0[i] + x2 * p0[i + 1])


def sl_dens_pot_force(r, lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale):
    """spheresl.get_halo_dens_pot_force (106-160) vectorised -> (dens, force, pot), each [l, n, N]."""
    x, dxi, i, x1, x2 = _sl_bins(r, xi, cmap, scale)
    ef = eftable[:lmax + 1, :nmax]
    sq = np.sqrt(evtable[:lmax + 1, :nmax])[:, :, None]
    fac = d_xi_to_r(x, cmap, scale) / dxi
    j = np.where(i == 0, 1, i)                        # 150-153: i==0 uses nodes 0,1,2
    dens = (x1 * ef[:, :, i] + x2 * ef[:, :, i + 1]) * sq * (x1 * d0[i] + x2 * d0[i + 1])
    force = fac * ((x2 - 0.5) * ef[:, :, j - 1] * p0[j - 1] - 2.0 * x2 * ef[:, :, j] * p0[j]
                   + (x2 + 0.5) * ef[:, :, j + 1] * p0[j + 1]) / sq
    pot = (x1 * ef[:, :, i] + x2 * ef[:, :, i + 1]) / sq * (x1 * p0[i] + x2 * p0[i + 1])
    return dens, force, pot


# ---------------------------------------------------------------------------
# SL accumulate -- spheresl.py:567-656
# ---------------------------------------------------------------------------
def sl_accumulate(x, y, z, mass, lmax, nmax, evtable, eftable, xi, p0, cmap, scale,
                  no_odd=False, chunk=16384):
    """
    expcoef[l^2 (m=0) | l^2+2m-1 (cos) | l^2+2m (sin), n]
      += -4pi m_p f[l,m] P_l^m(cos th) potd[l,n] {1 | cos m phi | sin m phi}
    with r = max(sqrt(r2), 1e-10) (611); no_odd skips odd l (630-632).
    """
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    z = np.asarray(z, np.float64); mass = np.asarray(mass, np.float64)
    expcoef = np.zeros((lmax * (lmax + 2) + 1, nmax))
    factorial = factorial_return(lmax)
    for lo in range(0, x.size, chunk):
        sl = slice(lo, lo + chunk)
        r2 = x[sl] * x[sl] + y[sl] * y[sl] + z[sl] * z[sl]
        r = np.fmax(np.sqrt(r2), 1.0e-10)
        costh = z[sl] / r
        phi = np.arctan2(y[sl], x[sl])
        legs = legendre_R(lmax, costh)
        potd = sl_pot_matrix(r, lmax, nmax, evtable, eftable, xi, p0, cmap, scale)
        loffset = 0
        for l in range(lmax + 1):
            if (l % 2 != 0) and no_odd:
                loffset += 2 * l + 1
                continue
            moffset = 0
            for m in range(l + 1):
                fac = factorial[l, m] * legs[l, m]
                fac4 = potd[l] * fac * FOURPI_NEG            # (nmax, N)
                if m == 0:
                    expcoef[loffset + moffset] += np.sum(fac4 * mass[sl], axis=1)
                    moffset += 1
                else:
                    expcoef[loffset + moffset] += np.sum(np.cos(phi * m) * fac4 * mass[sl], axis=1)
                    expcoef[loffset + moffset + 1] += np.sum(np.sin(phi * m) * fac4 * mass[sl], axis=1)
                    moffset += 2
            loffset += 2 * l + 1
    return expcoef


# ---------------------------------------------------------------------------
# SL field evaluation -- spheresl.py:987-1102 / 1107-1234 / 1240-1362
# ---------------------------------------------------------------------------
def _sl_field_sums(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                   evtable, eftable, l_lo, l_hi, no_odd, trig_index_l):
    """Shared sums for the three SL evaluation entry points (pot only; density out of scope)."""
    factorial = factorial_return(lmax)
    dend, dpot, potd = sl_dens_pot_force(r, lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale)
    legs, dlegs = dlegendre_R(lmax, costh)
    c = expcoef[:, :nmax]
    f00 = factorial[0, 0]
    pot0 = np.sum(f00 * c[0][:, None] * potd[0], axis=0)
    potr = np.sum(f00 * c[0][:, None] * dpot[0], axis=0)
    pot1 = np.zeros_like(pot0); pott = np.zeros_like(pot0); potp = np.zeros_like(pot0)
    loffset = 1
    for l in range(1, lmax + 1):
        if (l > l_hi) or (l < l_lo) or ((l % 2 != 0) and no_odd):
            loffset += 2 * l + 1
            continue
        moffset = 0
        for m in range(l + 1):
            f = factorial[l, m]
            if m == 0:
                s_p = np.sum(c[loffset + moffset][:, None] * potd[l], axis=0)
                s_d = np.sum(c[loffset + moffset][:, None] * dpot[l], axis=0)
                pot1 += f * legs[l, m] * s_p
                potr += f * legs[l, m] * s_d
                pott += f * dlegs[l, m] * s_p
                moffset += 1
            else:
                mt = l if trig_index_l else m         # spheresl.py:1173,1222-1225 quirk
                cosm = np.cos(phi * mt); sinm = np.sin(phi * mt)
                cc = c[loffset + moffset][:, None]; cs = c[loffset + moffset + 1][:, None]
                A_p = np.sum(cc * potd[l], axis=0) * cosm + np.sum(cs * potd[l], axis=0) * sinm
                A_d = np.sum(cc * dpot[l], axis=0) * cosm + np.sum(cs * dpot[l], axis=0) * sinm
                B_p = -np.sum(cc * potd[l], axis=0) * sinm + np.sum(cs * potd[l], axis=0) * cosm
                pot1 += f * legs[l, m] * A_p
                potr += f * legs[l, m] * A_d
                pott += f * dlegs[l, m] * A_p
                potp += f * legs[l, m] * m * B_p
                moffset += 2
        loffset += 2 * l + 1
    return pot0, pot1, potr, pott, potp


def _sl_density_sums(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                     evtable, eftable, l_lo, l_hi, no_odd, particles_variant):
    """
    Density outputs of the SL evaluation, quirks per function:
      all_eval (spheresl.py:1041-1092): den1 starts from the monopole (1046), m=0 term uses
        legs[l][0] (1071), densfac = 0.25/pi (1092);
      all_eval_particles (1289-1351): den1 starts from 0 (1271), its m=0 term uses legs[1][0]
        (1323), densfac = 0.25*pi (1351).
    Returns (den0, den1).
    """
    factorial = factorial_return(lmax)
    dend, _dpot, _potd = sl_dens_pot_force(r, lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale)
    legs = legendre_R(lmax, costh)
    c = expcoef[:, :nmax]
    den0 = np.sum(factorial[0, 0] * c[0][:, None] * dend[0], axis=0)
    den1 = np.zeros_like(den0) if particles_variant else den0.copy()
    loffset = 1
    for l in range(1, lmax + 1):
        if (l > l_hi) or (l < l_lo) or ((l % 2 != 0) and no_odd):
            loffset += 2 * l + 1
            continue
        moffset = 0
        for m in range(l + 1):
            f = factorial[l, m]
            if m == 0:
                leg = legs[1, 0] if particles_variant else legs[l, 0]
                den1 += f * leg * np.sum(c[loffset + moffset][:, None] * dend[l], axis=0)
                moffset += 1
            else:
                cosm = np.cos(phi * m); sinm = np.sin(phi * m)
                cc = c[loffset + moffset][:, None]; cs = c[loffset + moffset + 1][:, None]
                den1 += f * legs[l, m] * (np.sum(cc * dend[l], axis=0) * cosm + np.sum(cs * dend[l], axis=0) * sinm)
                moffset += 2
        loffset += 2 * l + 1
    densfac = 0.25 * np.pi if particles_variant else 0.25 / np.pi
    return den0 * densfac, den1 * densfac


def sl_all_eval_particles(x, y, z, expcoef, lmax, nmax, evtable, eftable, xi, p0, d0,
                          cmap, scale, L1=-1000, L2=1000, NO_ODD=False, chunk=16384, density=False):
    """
    spheresl.all_eval_particles (1240-1362): pot0, pot1, potr, pott, potp, rr with
    r = sqrt(x^2+y^2+z^2) (no epsilon, 1257), trig cos/sin(m phi).  With density=True the
    tuple is the reference's own (den0, den1, pot0, pot1, potr, pott, potp, rr).
    """
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64); z = np.asarray(z, np.float64)
    n = x.size
    out = [np.zeros(n) for _ in range(5)]
    den = [np.zeros(n) for _ in range(2)]
    rr = (x * x + y * y) ** 0.5
    for lo in range(0, n, chunk):
        sl = slice(lo, lo + chunk)
        r = (x[sl] * x[sl] + y[sl] * y[sl] + z[sl] * z[sl]) ** 0.5
        with np.errstate(invalid='ignore', divide='ignore'):
            costh = z[sl] / r
        phi = np.arctan2(y[sl], x[sl])
        res = _sl_field_sums(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                             evtable, eftable, L1, L2, NO_ODD, False)
        for o, v in zip(out, res):
            o[sl] = v
        if density:
            dres = _sl_density_sums(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                                    evtable, eftable, L1, L2, NO_ODD, True)
            for o, v in zip(den, dres):
                o[sl] = v
    pot0, pot1, potr, pott, potp = out
    if density:
        return den[0], den[1], pot0, pot1, potr, pott, potp, rr
    return pot0, pot1, potr, pott, potp, rr


def sl_force_eval(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                  evtable, eftable, no_odd=False):
    """
    spheresl.force_eval (1107-1234), vectorised over points.  lmax/nmax truncate
    (1138-1148); trig factors are cos/sin(l phi) (1173,1222-1225).  Returns
    (potr, pott, potp, pot1, pot0).
    """
    r = np.atleast_1d(np.asarray(r, np.float64)); costh = np.atleast_1d(np.asarray(costh, np.float64))
    phi = np.atleast_1d(np.asarray(phi, np.float64))
    ev = evtable[:lmax + 1, :nmax]; ef = eftable[:lmax + 1, :nmax]
    c = expcoef[:(lmax + 1) * (lmax + 1), :nmax]
    pot0, pot1, potr, pott, potp = _sl_field_sums(r, costh, phi, c, xi, p0, d0, cmap, scale,
                                                  lmax, nmax, ev, ef, -1000, 1000, no_odd, True)
    return potr, pott, potp, pot1, pot0


def sl_all_eval(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                evtable, eftable, no_odd=False, density=False):
    """spheresl.all_eval (987-1102) for full lmax/nmax: (pot0,pot1,potr,pott,potp), or with density=True
    the reference's own (den0,den1,pot0,pot1,potr,pott,potp)."""
    r = np.atleast_1d(np.asarray(r, np.float64)); costh = np.atleast_1d(np.asarray(costh, np.float64))
    phi = np.atleast_1d(np.asarray(phi, np.float64))
    pots = _sl_field_sums(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                          evtable, eftable, -1000, 1000, no_odd, False)
    if not density:
        return pots
    dens = _sl_density_sums(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax,
                            evtable, eftable, -1000, 1000, no_odd, False)
    return dens + tuple(pots)


# ---------------------------------------------------------------------------
# Fields.return_forces_cart -- potential.py:445-497
# ---------------------------------------------------------------------------
class FrozenField(object):
    """Plain container for what Fields.prep_tables (potential.py:257-294) loads."""

    def __init__(self, **kw):
        self.__dict__.update(kw)
        self.no_odd = False
        self.disk_use_m = self.mmax
        self.disk_use_n = self.norder
        self.halo_use_l = self.lmaxhalo
        self.halo_use_n = self.nmaxhalo

    def set_field_parameters(self, no_odd=False, halo_l=-1, halo_n=-1, disk_m=-1, disk_n=-1):
        """potential.py:365-377"""
        self.no_odd = no_odd
        if halo_l > -1: self.halo_use_l = halo_l
        if halo_n > -1: self.halo_use_n = halo_n
        if disk_m > -1: self.disk_use_m = disk_m
        if disk_n > -1: self.disk_use_n = disk_n


def fields_forces_cart(F, xval, yval, zval, rotpos=0.0):
    """
    Fields.return_forces_cart (potential.py:445-497), vectorised over points.
    F: FrozenField.  Returns the 8-tuple
    (fxdisk, fxhalo, fydisk, fyhalo, fzdisk, fzhalo, diskp, halop+halop0).
    """
    xval = np.atleast_1d(np.asarray(xval, np.float64)); yval = np.atleast_1d(np.asarray(yval, np.float64))
    zval = np.atleast_1d(np.asarray(zval, np.float64))
    r2val = (xval * xval + yval * yval) ** 0.5 + 1.e-15
    r3val = (r2val * r2val + zval * zval) ** 0.5 + 1.e-15
    costh = zval / r3val
    phival = np.arctan2(yval, xval)
    diskfr, diskfp, diskfz, diskp, diskp0 = eof_force_eval(
        r2val, zval, phival + rotpos, F.cos, F.sin, F.potC, F.rforceC, F.zforceC,
        F.potS, F.rforceS, F.zforceS, F.XMIN, F.dX, F.YMIN, F.dY, F.numx, F.numy,
        F.disk_use_m, F.disk_use_n, F.ascale, F.hscale, F.cmapdisk, no_odd=F.no_odd)
    halofr, haloft, halofp, halop, halop0 = sl_force_eval(
        r3val, costh, phival + rotpos, F.halofac * F.expcoef, F.xihalo, F.p0halo, F.d0halo,
        F.cmaphalo, F.scalehalo, F.halo_use_l, F.halo_use_n, F.evtablehalo, F.eftablehalo,
        no_odd=F.no_odd)
    guard = r3val < np.min(F.xihalo)                       # potential.py:483-485
    halofp = np.where(guard, 0., halofp)
    diskfp = np.where(guard, 0., diskfp)
    fxdisk = (diskfr * (xval / r2val) - diskfp * (yval / (r2val * r2val)))
    fxhalo = -1. * (halofr * (xval / r3val) - haloft * (xval * zval / (r3val * r3val * r3val))) + halofp * (yval / (r2val * r2val))
    fydisk = (diskfr * (yval / r2val) + diskfp * (xval / (r2val * r2val)))
    fyhalo = -1. * (halofr * (yval / r3val) - haloft * (yval * zval / (r3val * r3val * r3val))) - halofp * (xval / (r2val * r2val))
    fzdisk = diskfz
    fzhalo = -1. * (halofr * (zval / r3val) + haloft * ((r2val * r2val) / (r3val * r3val * r3val)))
    return fxdisk, fxhalo, fydisk, fyhalo, fzdisk, fzhalo, diskp, (halop + halop0)


def fields_forces_cyl(F, xval, yval, zval, rotpos=0.0):
    """
    Fields.return_forces_cyl (potential.py:389-440), vectorised over points:
    (diskfr, frhalo, diskfp, -halofp, diskfz, fzhalo, -diskp, halop+halop0); r2, r3 use +1e-10.
    """
    xval = np.atleast_1d(np.asarray(xval, np.float64)); yval = np.atleast_1d(np.asarray(yval, np.float64))
    zval = np.atleast_1d(np.asarray(zval, np.float64))
    r2val = np.sqrt(xval * xval + yval * yval) + 1.e-10
    r3val = np.sqrt(r2val * r2val + zval * zval) + 1.e-10
    costh = zval / r3val
    phival = np.arctan2(yval, xval)
    diskfr, diskfp, diskfz, diskp, diskp0 = eof_force_eval(
        r2val, zval, phival + rotpos, F.cos, F.sin, F.potC, F.rforceC, F.zforceC,
        F.potS, F.rforceS, F.zforceS, F.XMIN, F.dX, F.YMIN, F.dY, F.numx, F.numy,
        F.disk_use_m, F.disk_use_n, F.ascale, F.hscale, F.cmapdisk, no_odd=F.no_odd)
    halofr, haloft, halofp, halop, halop0 = sl_force_eval(
        r3val, costh, phival + rotpos, F.halofac * F.expcoef, F.xihalo, F.p0halo, F.d0halo,
        F.cmaphalo, F.scalehalo, F.halo_use_l, F.halo_use_n, F.evtablehalo, F.eftablehalo,
        no_odd=F.no_odd)
    guard = r3val < np.min(F.xihalo)                       # potential.py:427-429
    halofp = np.where(guard, 0., halofp)
    diskfp = np.where(guard, 0., diskfp)
    frhalo = -1. * (r2val * halofr + zval * haloft) / r3val
    fzhalo = -1. * (zval * halofr - r2val * haloft) / r3val
    return diskfr, frhalo, diskfp, -1. * halofp, diskfz, fzhalo, -1. * diskp, (halop + halop0)


# ---------------------------------------------------------------------------
# leapfrog -- integrate.py:53-190
# ---------------------------------------------------------------------------
def leapfrog(F, nint, dt, pos0, vel0, rotfreq=0., no_odd=False, halo_l=-1, halo_n=-1,
             disk_m=-1, disk_n=-1, keep_trajectory=False):
    """
    integrate.leapfrog_integrate (53-190) vectorised over orbits (apse=False):
    pos0, vel0 shape (3, norb).  barpos_k = 2 pi rotfreq k dt (97);
    x1 = x0 + v0 dt + 0.5 a0 dt^2 (129-131); v1 = v0 + 0.5 (a0+a1) dt (141-143).
    Returns end state (pos, vel, pot) at step nint-1, plus the full trajectory
    dict (T,X,Y,Z,VX,VY,VZ,P, each (nint, norb)) if keep_trajectory.
    """
    F.set_field_parameters(no_odd=no_odd, halo_l=halo_l, halo_n=halo_n, disk_m=disk_m, disk_n=disk_n)
    pos = np.array(pos0, dtype=np.float64).reshape(3, -1).copy()
    vel = np.array(vel0, dtype=np.float64).reshape(3, -1).copy()
    times = np.arange(0, nint, 1) * dt
    barpos = 2. * np.pi * rotfreq * times
    traj = None
    if keep_trajectory:
        norb = pos.shape[1]
        traj = {k: np.zeros((nint, norb)) for k in ('X', 'Y', 'Z', 'VX', 'VY', 'VZ', 'P', 'FX', 'FY', 'FZ')}
        traj['T'] = times

    def force(p, rot):
        dfx, hfx, dfy, hfy, dfz, hfz, dp, hp = fields_forces_cart(F, p[0], p[1], p[2], rotpos=rot)
        return np.stack([dfx + hfx, dfy + hfy, dfz + hfz]), dp + hp

    a0, pot = force(pos, barpos[0])

    def record(k):
        if traj is not None:
            traj['X'][k], traj['Y'][k], traj['Z'][k] = pos
            traj['VX'][k], traj['VY'][k], traj['VZ'][k] = vel
            traj['FX'][k], traj['FY'][k], traj['FZ'][k] = a0
            traj['P'][k] = pot
    record(0)
    for step in range(1, nint):
        pos = pos + (vel * dt) + (0.5 * a0 * (dt ** 2.))
        a1, pot = force(pos, barpos[step])
        vel = vel + (0.5 * (a0 + a1) * dt)
        a0 = a1
        record(step)
    return pos, vel, pot, traj


# ---------------------------------------------------------------------------
# pre-accumulation transforms
# ---------------------------------------------------------------------------
def bar_fourier_angle(posx, posy, minr=0., maxr=1.):
    """pattern.BarTransform.bar_fourier_compute (analysis/pattern.py:155-169)."""
    rr = (posx * posx + posy * posy) ** 0.5
    w = np.where((rr > minr) & (rr < maxr))[0]
    aval = np.sum(np.cos(2. * np.arctan2(posy[w], posx[w])))
    bval = np.sum(np.sin(2. * np.arctan2(posy[w], posx[w])))
    return np.arctan2(bval, aval) / 2.


def bar_rotate(x, y, bar_angle):
    """pattern.BarTransform.calculate_transform_and_return (pattern.py:118-121)."""
    return x * np.cos(bar_angle) - y * np.sin(bar_angle), x * np.sin(bar_angle) + y * np.cos(bar_angle)


def inner_center(xr, yr, zr, xv, yv, zv, mv, ncenter=10000):
    """Fields.total_coefficients centring (potential.py:158-176, 190-200): rank by (x^2+y^2+z^2)^0.5 of the
    r-arrays, mass-weighted mean of the v-arrays over the first ncenter indices."""
    rrank = (xr * xr + yr * yr + zr * zr) ** 0.5
    c = rrank.argsort()[0:ncenter]
    return (np.sum(xv[c] * mv[c]) / np.sum(mv[c]), np.sum(yv[c] * mv[c]) / np.sum(mv[c]),
            np.sum(zv[c] * mv[c]) / np.sum(mv[c]))
