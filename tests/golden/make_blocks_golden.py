"""
make_blocks_golden.py -- golden vectors for the per-point building blocks (SURVEY.md section 8a rows a5, a6,
a11 accumulated_eval, a14, a15, a16), produced by the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_blocks_golden.py

Writes tests/golden/blocks_small.npz (random tables, cmap 1 for both bases, dens=1 EOF cache).
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as G                          # noqa: E402  (reference modules + fixture writers)
from exptool_b200 import synthetic as S          # noqa: E402

eof, spheresl, halo_methods = G.eof, G.spheresl, G.halo_methods


def main():
    seed = 61
    eparams = dict(G.SMALL_EOF, cmap=1, dens=1)
    sparams = dict(G.SMALL_SL, cmap=1)
    rng = np.random.default_rng(seed + 100)
    with tempfile.TemporaryDirectory() as tmp:
        f, tabs, g = G.eof_setup(tmp, eparams, 'random', seed)
        potC, rforceC, zforceC, densC, potS, rforceS, zforceS, densS = tabs
        pe, _ = S.make_eof_tables(eparams, kind='random', seed=seed)
        n = 40
        x, y, z, _m = G.edge_particles(rng, n, 3 * pe['ascale'], 3 * pe['hscale'])
        r = np.sqrt(x * x + y * y + 1e-10)
        phi = np.arctan2(y, x)
        geo = dict(rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'], numy=g['numy'],
                   ASCALE=g['ascale'], HSCALE=g['hscale'], CMAP=g['cmap'])
        # eof.return_bins (eof.py:354-427), vector and scalar call
        X, Y, ix, iy = eof.return_bins(r.copy(), z.copy(), **geo)
        Xs, Ys, ixs, iys = eof.return_bins(float(r[9]), float(z[9]), **geo)
        # eof.get_pot (eof.py:430-457)
        Vc, Vs = eof.get_pot(r.copy(), z.copy(), potC, potS, fac=1.0, MMAX=g['mmax'], NMAX=g['norder'], **geo)
        # eof.accumulated_eval (eof.py:874-929), point by point
        cosc = rng.normal(0, 1, (g['mmax'] + 1, g['norder'])); sinc = rng.normal(0, 1, (g['mmax'] + 1, g['norder']))
        sinc[0] = 0.0
        ae, ae_noodd = [], []
        for i in range(n):
            a = (float(r[i]), float(z[i]), float(phi[i]), cosc, sinc, potC, rforceC, zforceC, densC, potS, rforceS,
                 zforceS, densS)
            kw = dict(geo, MMAX=g['mmax'], NMAX=g['norder'])
            ae.append(eof.accumulated_eval(*a, **kw))
            ae_noodd.append(eof.accumulated_eval(*a, no_odd=True, **kw))
        # SL radial matrices and Legendre tables
        sf, mf, ps = G.sl_setup(tmp, sparams, 'random', seed + 1)
        lmax, nmax, numr, cmap, rmin, rmax, scale, ltable, evtable, eftable = halo_methods.read_cached_table(sf)
        xi, rarr, p0, d0 = G.quiet(halo_methods.init_table, mf, numr, rmin, rmax, cmap=cmap, scale=scale)
        rad = np.concatenate([[1e-12, 1e-5, rmin, 0.5 * rmin, rmax, 3 * rmax, 40.0],
                              scale * 10 ** rng.uniform(-2.5, 1.4, 25)])
        dens, force, pot, potm, densm = [], [], [], [], []
        for rv in rad:
            dd, ff, pp = spheresl.get_halo_dens_pot_force(float(rv), lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale)
            dens.append(dd); force.append(ff); pot.append(pp)
            potm.append(spheresl.get_halo_pot_matrix(float(rv), lmax, nmax, evtable, eftable, xi, p0, cmap, scale))
            densm.append(spheresl.get_halo_dens(float(rv), lmax, nmax, evtable, eftable, xi, d0, cmap, scale))
        cth = np.concatenate([[1.0, -1.0, 0.0, 1.0 - 1e-9, -1.0 + 1e-12, 0.999999], rng.uniform(-1, 1, 26)])
        L = 6
        P, P2, dP = [], [], []
        for c in cth:
            P.append(spheresl.legendre_R(L, float(c)))
            a, b = spheresl.dlegendre_R(L, float(c))
            P2.append(a); dP.append(b)
    np.savez_compressed(os.path.join(HERE, 'blocks_small.npz'),
                        meta=json.dumps(dict(eof_params=eparams, sl_params=sparams, kind='random', seed=seed, geo=g,
                                             lmax=int(lmax), nmax=int(nmax), leg_lmax=L)),
                        r=r, z=z, phi=phi, X=X, Y=Y, ix=ix, iy=iy,
                        bins_scalar=np.array([float(Xs), float(Ys), float(ixs), float(iys)]),
                        Vc=Vc, Vs=Vs, cosc=cosc, sinc=sinc, ae=np.array(ae, dtype=np.float64),
                        ae_noodd=np.array(ae_noodd, dtype=np.float64),
                        rad=rad, dens=np.array(dens), force=np.array(force), pot=np.array(pot), potm=np.array(potm),
                        densm=np.array(densm), cth=cth, P=np.array(P), P2=np.array(P2), dP=np.array(dP))
    print('wrote blocks_small')


if __name__ == '__main__':
    main()
