"""
synthetic.py -- synthetic basis tables and particle sets for tests and bench.

exptool never builds basis tables itself (eof.py:92-96) and ships no cache
file, so every hot-path function needs synthesized inputs.
This module writes them in the reference's own on-disk formats so that both
the reference readers and ours consume identical bytes:

  * EOF cache, old style        -> read by eof.eof_params / eof.parse_eof
                                   (eof.py:159-181, 257-308)
  * SL cache                    -> read by halo_methods.read_cached_table
                                   (halo_methods.py:128-164)
  * spherical model file R D M P-> read by halo_methods.read_sph_model_table
                                   (halo_methods.py:56-62)

and samples the Hernquist halo / exponential disc particle sets of
SURVEY.md section 8d.  Everything here is host-side NumPy; nothing is on the timed path.
"""
import os
import numpy as np

# ---------------------------------------------------------------------------
# geometry defaults (the shipped run headers / eof.set_table_params defaults)
# ---------------------------------------------------------------------------
EOF_DEFAULT = dict(mmax=6, numx=128, numy=64, nmax=64, norder=18, dens=0, cmap=1,
                   rmin=0.001, rmax=20.0, ascale=0.01, hscale=0.001,
                   cylmass=1.0, time=0.0)
SL_DEFAULT = dict(lmax=4, nmax=18, numr=2000, cmap=1, rmin=1.0e-4, rmax=1.95,
                  scale=0.0667)


def _r_to_xi(r, cmap, scale):
    if cmap == 1:
        return (r / scale - 1.0) / (r / scale + 1.0)
    if cmap == 2:
        return np.log(r)
    return r


def _xi_to_r(xi, cmap, scale):
    if cmap == 1:
        return (1.0 + xi) / (1.0 - xi) * scale
    if cmap == 2:
        return np.exp(xi)
    return xi


def eof_node_coordinates(p):
    """(R_i, z_j) of the (numx+1) x (numy+1) table nodes for header dict p."""
    rtable = np.sqrt(0.5) * p['rmax']
    xmin = _r_to_xi(p['rmin'] * p['ascale'], p['cmap'], p['ascale'])
    xmax = _r_to_xi(rtable * p['ascale'], p['cmap'], p['ascale'])
    ymax = np.arcsinh(rtable * p['ascale'] / p['hscale'])
    xi = np.linspace(xmin, xmax, p['numx'] + 1)
    yy = np.linspace(-ymax, ymax, p['numy'] + 1)
    return _xi_to_r(xi, p['cmap'], p['ascale']), p['hscale'] * np.sinh(yy)


def make_eof_tables(params=None, kind='smooth', seed=0):
    """
    Build the six (+2 density) EOF tables, shape (mmax+1, norder, numx+1, numy+1).

    kind='smooth' : analytic family  pot_mn = -g_n (rho/alpha_n)^m q^-(m+1)/2,
                    q = 1 + rho^2/alpha_n^2 + z^2/b_n^2, with rforce = -dpot/dR
                    and zforce = -dpot/dz exact, so that orbits in the summed
                    field are bound and smooth.  Sine tables equal the cosine
                    tables for m>=1 and are zero for m=0 (as parse_eof leaves them).
    kind='random' : standard_normal tables (adversarial parity fixture).
    """
    p = dict(EOF_DEFAULT)
    if params:
        p.update(params)
    M, N, NX, NY = p['mmax'] + 1, p['norder'], p['numx'] + 1, p['numy'] + 1
    shape = (M, N, NX, NY)
    if kind == 'random':
        rng = np.random.default_rng(seed)
        T = {k: rng.standard_normal(shape) for k in
             ('potC', 'rforceC', 'zforceC', 'potS', 'rforceS', 'zforceS')}
        for k in ('potS', 'rforceS', 'zforceS'):
            T[k][0] = 0.0
        T['densC'] = rng.standard_normal(shape) if p['dens'] else np.zeros(shape)
        T['densS'] = rng.standard_normal(shape) if p['dens'] else np.zeros(shape)
        T['densS'][0] = 0.0
        return p, T

    R, z = eof_node_coordinates(p)
    a = p['ascale']
    rho = (R / a)[:, None]
    zz = z[None, :]
    potC = np.zeros(shape)
    rfC = np.zeros(shape)
    zfC = np.zeros(shape)
    for m in range(M):
        for n in range(N):
            alpha = 1.0 + 0.5 * n
            b = a * (0.2 + 0.1 * n)
            g = 1.0 / (1.0 + n) * (-1.0) ** n if n > 0 else 1.0
            q = 1.0 + rho ** 2 / alpha ** 2 + zz ** 2 / b ** 2
            e1 = -0.5 * (m + 1)
            rm = (rho / alpha) ** m
            pot = -g * rm * q ** e1
            drm = m * (rho / alpha) ** (m - 1) / alpha if m > 0 else np.zeros_like(rho)
            dpot_dR = -g * (drm * q ** e1 + rm * e1 * q ** (e1 - 1.0) * 2.0 * rho / alpha ** 2) / a
            dpot_dz = -g * (rm * e1 * q ** (e1 - 1.0) * 2.0 * zz / b ** 2)
            potC[m, n] = pot
            rfC[m, n] = -dpot_dR
            zfC[m, n] = -dpot_dz
    T = dict(potC=potC, rforceC=rfC, zforceC=zfC,
             potS=potC.copy(), rforceS=rfC.copy(), zforceS=zfC.copy())
    for k in ('potS', 'rforceS', 'zforceS'):
        T[k][0] = 0.0
    T['densC'] = np.zeros(shape)
    T['densS'] = np.zeros(shape)
    if p['dens']:
        T['densC'] = -potC * 0.25 / np.pi
        T['densS'] = -T['potS'] * 0.25 / np.pi
    return p, T


def write_eof_cache(path, p, T):
    """Old-style EOF cache, byte layout of eof.py:159-181 (header) and 274-308 (body)."""
    with open(path, 'wb') as f:
        np.array([p['mmax'], p['numx'], p['numy'], p['nmax'], p['norder'],
                  p['dens'], p['cmap']], dtype='<u4').tofile(f)
        np.array([p['rmin'], p['rmax'], p['ascale'], p['hscale'],
                  p['cylmass'], p['time']], dtype='<f8').tofile(f)
        for m in range(p['mmax'] + 1):
            for n in range(p['norder']):
                T['potC'][m, n].astype('<f8').tofile(f)
                T['rforceC'][m, n].astype('<f8').tofile(f)
                T['zforceC'][m, n].astype('<f8').tofile(f)
                if p['dens']:
                    T['densC'][m, n].astype('<f8').tofile(f)
        for m in range(1, p['mmax'] + 1):
            for n in range(p['norder']):
                T['potS'][m, n].astype('<f8').tofile(f)
                T['rforceS'][m, n].astype('<f8').tofile(f)
                T['zforceS'][m, n].astype('<f8').tofile(f)
                if p['dens']:
                    T['densS'][m, n].astype('<f8').tofile(f)
    return path


def make_sl_tables(params=None, kind='smooth', seed=0):
    """evtable (lmax+1,nmax) > 0 increasing in n; eftable (lmax+1,nmax,numr)."""
    p = dict(SL_DEFAULT)
    if params:
        p.update(params)
    L, N, NR = p['lmax'] + 1, p['nmax'], p['numr']
    if kind == 'random':
        rng = np.random.default_rng(seed)
        ev = 0.5 + rng.random((L, N)) + np.arange(N)[None, :]
        ef = rng.standard_normal((L, N, NR))
        return p, ev, ef
    ximin = _r_to_xi(p['rmin'], p['cmap'], p['scale'])
    ximax = _r_to_xi(p['rmax'], p['cmap'], p['scale'])
    xi = np.linspace(ximin, ximax, NR)
    u = (xi - ximin) / (ximax - ximin)
    ev = np.zeros((L, N))
    ef = np.zeros((L, N, NR))
    for l in range(L):
        for n in range(N):
            ev[l, n] = (n + 1.0 + 0.5 * l) ** 2
            ef[l, n] = np.cos(n * np.pi * u + 0.3 * l) * (0.5 + 0.5 / (1.0 + l))
    return p, ev, ef


def write_sl_cache(path, p, ev, ef):
    """SL cache, byte layout of halo_methods.py:128-164."""
    with open(path, 'wb') as f:
        np.array([p['lmax'], p['nmax'], p['numr'], p['cmap']], dtype='<u4').tofile(f)
        np.array([p['rmin'], p['rmax'], p['scale']], dtype='<f8').tofile(f)
        for l in range(p['lmax'] + 1):
            np.array([l], dtype='<u4').tofile(f)
            ev[l].astype('<f8').tofile(f)
            for n in range(p['nmax']):
                ef[l, n].astype('<f8').tofile(f)
    return path


def write_hernquist_model(path, a=0.0667, mass=1.0, rmin=5.0e-5, rmax=3.0, nr=1000):
    """
    Text model file 'R D M P' (writer format models/twopower.py:186-199; reader
    halo_methods.py:56-62 skips 5 lines and '!' comments).  Analytic Hernquist
    (models/hernquist.py:30-73): rho = M a / (2 pi r (r+a)^3), M(r) = M r^2/(r+a)^2,
    Phi = -M/(r+a).
    """
    r = np.logspace(np.log10(rmin), np.log10(rmax), nr)
    d = mass * a / (2.0 * np.pi * r * (r + a) ** 3)
    m = mass * r ** 2 / (r + a) ** 2
    ph = -mass / (r + a)
    with open(path, 'w') as f:
        f.write('! Hernquist a=%g M=%g (synthetic)\n' % (a, mass))
        f.write('! R    D    M    P\n')
        f.write('! written by exptool_b200.synthetic\n')
        f.write('! first five lines are skipped by the reader\n')
        f.write('%d\n' % nr)
        for i in range(nr):
            f.write('%.16e %.16e %.16e %.16e\n' % (r[i], d[i], m[i], ph[i]))
    return path


# ---------------------------------------------------------------------------
# particle samplers
# ---------------------------------------------------------------------------
def hernquist_halo(n, seed, a=0.0667, rmax=1.95):
    """Equal-mass Hernquist sphere, inverse CDF of M(r)=r^2/(r+a)^2 truncated at rmax."""
    rng = np.random.default_rng(seed)
    mmax = rmax ** 2 / (rmax + a) ** 2
    u = rng.random(n) * mmax
    s = np.sqrt(u)
    r = a * s / (1.0 - s)
    cth = rng.uniform(-1.0, 1.0, n)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    sth = np.sqrt(1.0 - cth * cth)
    x = r * sth * np.cos(phi)
    y = r * sth * np.sin(phi)
    z = r * cth
    m = np.full(n, 1.0 / n)
    return x, y, z, m


def exponential_disc(n, seed, a=0.01, h=0.001, rtrunc=0.14):
    """Sigma ~ exp(-R/a) (R ~ Gamma(2,a)), sech^2 layer z = h atanh(U(-1,1)), R < rtrunc."""
    rng = np.random.default_rng(seed)
    R = rng.gamma(2.0, a, n)
    bad = R >= rtrunc
    while bad.any():
        R[bad] = rng.gamma(2.0, a, int(bad.sum()))
        bad = R >= rtrunc
    u = rng.uniform(-1.0, 1.0, n)
    u = np.clip(u, -1.0 + 1e-12, 1.0 - 1e-12)
    z = h * np.arctanh(u)
    phi = rng.uniform(0.0, 2.0 * np.pi, n)
    x = R * np.cos(phi)
    y = R * np.sin(phi)
    m = np.full(n, 1.0 / n)
    return x, y, z, m


class ParticleSet(object):
    """Minimal object with the `.data` dict interface eof/spheresl accept
    (eof.py:582; also exposes holder-style attributes, which
    eof.accumulated_eval_particles needs at eof.py:1079)."""

    def __init__(self, x, y, z, m, time=0.0, filename='synthetic', comp='synthetic'):
        self.data = {'x': x, 'y': y, 'z': z, 'm': m}
        self.xpos, self.ypos, self.zpos, self.mass = x, y, z, m
        self.time = time
        self.filename = filename
        self.comp = comp


def write_fixture_files(dirname, eof_params=None, sl_params=None, kind='smooth', seed=0):
    """Write eof cache, sl cache, model file into dirname; return their paths."""
    os.makedirs(dirname, exist_ok=True)
    pe, T = make_eof_tables(eof_params, kind=kind, seed=seed)
    ps, ev, ef = make_sl_tables(sl_params, kind=kind, seed=seed + 1)
    tag = '%s_m%d_n%d_l%d' % (kind, pe['mmax'], pe['norder'], ps['lmax'])
    eof_file = write_eof_cache(os.path.join(dirname, '.eof.cache.' + tag), pe, T)
    sl_file = write_sl_cache(os.path.join(dirname, 'SLGridSph.cache.' + tag), ps, ev, ef)
    model_file = write_hernquist_model(os.path.join(dirname, 'SLGridSph.model'), a=ps['scale'])
    return eof_file, sl_file, model_file


def barred_snapshot(seed, ndisc, nhalo, bar_angle=0.6, bar_strength=0.35, center=(0.0021, -0.0013, 0.0004),
                    unequal_mass=True):
    """
    Synthetic two-component snapshot for the ingest path (Fields.total_coefficients): an exponential disc with an
    m = 2 distortion (a fraction `bar_strength` of the particles is squeezed towards the axis at `bar_angle`)
    and a Hernquist halo, both displaced by `center`, with velocities and potential energies so that every PSP
    field is populated.  Returns {'star': data, 'dark': data} with data = dict(m,x,y,z,vx,vy,vz,potE).
    """
    rng = np.random.default_rng(seed)
    x, y, z, m = exponential_disc(ndisc, seed + 1)
    R = np.hypot(x, y); phi = np.arctan2(y, x)
    squeeze = rng.random(ndisc) < bar_strength
    phi = np.where(squeeze, bar_angle + 0.35 * (phi - bar_angle) * 0.5 + np.pi * rng.integers(0, 2, ndisc), phi)
    x, y = R * np.cos(phi), R * np.sin(phi)
    if unequal_mass:
        m = m * rng.uniform(0.5, 1.5, ndisc)
    vc = np.sqrt(R / (R + 0.01)) * 1.2
    out = {}
    out['star'] = dict(m=m, x=x + center[0], y=y + center[1], z=z + center[2],
                       vx=-vc * np.sin(phi) + rng.normal(0, 0.05, ndisc), vy=vc * np.cos(phi) + rng.normal(0, 0.05, ndisc),
                       vz=rng.normal(0, 0.02, ndisc), potE=-1.0 / (R + 0.01))
    xh, yh, zh, mh = hernquist_halo(nhalo, seed + 2)
    if unequal_mass:
        mh = mh * rng.uniform(0.5, 1.5, nhalo)
    rh = np.sqrt(xh * xh + yh * yh + zh * zh)
    out['dark'] = dict(m=mh, x=xh + center[0], y=yh + center[1], z=zh + center[2],
                       vx=rng.normal(0, 0.3, nhalo), vy=rng.normal(0, 0.3, nhalo), vz=rng.normal(0, 0.3, nhalo),
                       potE=-1.0 / (rh + 0.0667))
    return out
