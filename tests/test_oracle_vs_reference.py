"""
Live cross-check of the oracle against the UNMODIFIED reference on fresh seeds
(not the committed goldens).  Runs only where /root/reference exists (the build
container); skipped on the GPU box.  CPU only.
"""
import io
import os
import contextlib
import tempfile

import numpy as np
import pytest

from helpers import relerr, O, S
from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason='reference tree not present')


@pytest.fixture(scope='module')
def ref():
    return refshim.load()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def test_eof_accumulate_and_force_fresh_seed(ref):
    eof = ref['eof']
    with tempfile.TemporaryDirectory() as tmp:
        pe, T = S.make_eof_tables(dict(mmax=4, numx=24, numy=16, nmax=8, norder=5, cmap=1), kind='random', seed=77)
        f = S.write_eof_cache(os.path.join(tmp, 'c'), pe, T)
        tabs = quiet(eof.parse_eof, f)
        rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, dens = eof.eof_params(f)
        XMIN, XMAX, dX, YMIN, YMAX, dY = eof.set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ascale, HSCALE=hscale,
                                                              NUMX=numx, NUMY=numy, CMAP=cmap)
    x, y, z, m = S.exponential_disc(700, 4242)
    P = S.ParticleSet(x, y, z, m)
    c, s = eof.accumulate(P, tabs[0], tabs[4], mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap)
    geo = (float(XMIN), float(dX), float(YMIN), float(dY), int(numx), int(numy))
    co, so = O.eof_accumulate(x, y, z, m, T['potC'], T['potS'], int(mmax), int(norder), *geo, ascale, hscale, int(cmap))
    assert relerr(co, c) < 1e-12 and relerr(so, s) < 1e-12
    Pf = S.ParticleSet(x[:60], y[:60], z[:60], m[:60])
    ref_out = eof.accumulated_eval_particles(Pf, c, s, potC=tabs[0], rforceC=tabs[1], zforceC=tabs[2], potS=tabs[4],
                                             rforceS=tabs[5], zforceS=tabs[6], rmin=XMIN, dR=dX, zmin=YMIN, dZ=dY,
                                             numx=numx, numy=numy, MMAX=mmax, NMAX=norder, ASCALE=ascale,
                                             HSCALE=hscale, CMAP=cmap, verbose=0)
    out = O.eof_force_particles(x[:60], y[:60], z[:60], c, s, T['potC'], T['rforceC'], T['zforceC'], T['potS'],
                                T['rforceS'], T['zforceS'], *geo, int(mmax), int(norder), ascale, hscale, int(cmap))
    for i in range(6):
        assert relerr(out[i], ref_out[i]) < 1e-12, i


def test_sl_accumulate_fresh_seed(ref):
    spheresl, particle = ref['spheresl'], ref['particle']
    from helpers import hernquist_model_columns
    with tempfile.TemporaryDirectory() as tmp:
        ps, ev, ef = S.make_sl_tables(dict(lmax=3, nmax=4, numr=120), kind='random', seed=91)
        sf = S.write_sl_cache(os.path.join(tmp, 's'), ps, ev, ef)
        mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
        x, y, z, m = S.hernquist_halo(150, 515)
        H = particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = x, y, z, m
        c = np.asarray(quiet(spheresl.compute_coefficients_solitary, H, sf, mf), dtype=np.float64)
    R1, D1, P1 = hernquist_model_columns(ps['scale'])
    xi, r, p0, d0 = O.sl_init_table(R1, D1, P1, ps['numr'], ps['rmin'], ps['rmax'], ps['cmap'], ps['scale'])
    co = O.sl_accumulate(x, y, z, m, ps['lmax'], ps['nmax'], ev, ef, xi, p0, ps['cmap'], ps['scale'])
    assert relerr(co, c) < 1e-12


def test_factorial_return_and_mirrors_vs_reference(ref):
    """spheresl.factorial_return (spheresl.py:823-863), compatibility maps and eof.set_table_params of the product's
    host-side mirrors against the reference's own functions (host code: no GPU needed)."""
    from exptool_b200.basis import spheresl as mine, eof as myeof
    for lmax in (0, 2, 6, 9):
        r = ref['spheresl'].factorial_return(lmax)
        assert np.array_equal(np.asarray(r), mine.factorial_return(lmax))
        assert np.allclose(O.factorial_return(lmax), r, rtol=1e-15, atol=0)
    for cmap in (0, 1):
        a = ref['eof'].set_table_params(RMAX=20.0, RMIN=0.001, ASCALE=0.01, HSCALE=0.001, NUMX=128, NUMY=64, CMAP=cmap)
        b = myeof.set_table_params(RMAX=20.0, RMIN=0.001, ASCALE=0.01, HSCALE=0.001, NUMX=128, NUMY=64, CMAP=cmap)
        assert np.array_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64))


def test_leapfrog_foreign_field_and_bar_pattern_vs_reference(ref):
    """The host-side pieces of the workflow drivers against the reference's own code: leapfrog_integrate on a foreign
    (duck-typed) field object gives the reference's arrays bit for bit; BarDetermine.read_bar / frequency_and_derivative /
    find_barpattern give the reference's pattern speed from the same printed bar file."""
    from exptool_b200.utils import integrate as mine
    from exptool_b200.analysis import pattern as mypattern

    class Kepler(object):
        def set_field_parameters(self, **kw):
            pass

        def return_forces_cart(self, x, y, z, rotpos=0.0):
            r = np.sqrt(x * x + y * y + z * z)
            return -x / r ** 3, 0.0, -y / r ** 3, 0.0, -z / r ** 3, 0.0, -1.0 / r, 0.0
    a = (Kepler(), 300, 0.01, [1.0, 0.0, 0.02], [0.0, 0.8, 0.05])
    for kw in (dict(rotfreq=0.3, force=True), dict(rotfreq=-0.3, apse=True, ap_max=1), dict(ap_max=0)):
        Or = quiet(ref['integrate'].leapfrog_integrate, *a, **kw)
        Om = mine.leapfrog_integrate(*a, **kw)
        assert set(Or.keys()) == set(Om.keys())
        for k in Or.keys():
            assert np.array_equal(np.asarray(Or[k]), np.asarray(Om[k])), (kw, k)
    with tempfile.TemporaryDirectory() as tmp:
        bf = os.path.join(tmp, 'bar.dat')
        t = np.linspace(0.0, 2.0, 101)
        with open(bf, 'w') as f:
            for ti in t:
                f.write('%.6f %.8f\n' % (ti, 37.5 * ti - 3.0 * ti * ti))
        Br, Bm = ref['pattern'].BarDetermine(), mypattern.BarDetermine()
        Br.read_bar(bf); Bm.read_bar(bf)
        assert np.array_equal(Br.time, Bm.time) and np.array_equal(Br.deriv, Bm.deriv)
        Br.frequency_and_derivative(spline_derivative=2); Bm.frequency_and_derivative(spline_derivative=2)
        assert np.array_equal(Br.deriv, Bm.deriv) and np.array_equal(Br.dderiv, Bm.dderiv)
        for tt in (0.0, 0.777, np.array([0.1, 1.9])):
            assert np.array_equal(np.asarray(ref['pattern'].find_barpattern(tt, Br, smth_order=None)),
                                  np.asarray(mypattern.find_barpattern(tt, Bm, smth_order=None)))
