"""EXP `outcoef` wire format: exptool_b200.io.outcoef against the reference's own
fixture (exptool/tests/outcoef.star.run0.dat) as read by the unmodified reference reader, plus writer round trips."""
import contextlib
import hashlib
import io
import os

import numpy as np
import pytest

from helpers import GOLDEN
from exptool_b200.io import outcoef

FIXTURE = '/root/reference/exptool/tests/outcoef.star.run0.dat'


def test_reader_matches_reference_on_fixture_head():
    g = np.load(os.path.join(GOLDEN, 'outcoef_star.npz'))
    o = outcoef.OutCoef(os.path.join(GOLDEN, 'outcoef_star_head.dat'), verbose=0)
    assert o.basis == 'Cylinder'
    assert np.array_equal(o.T, g['T_head'])
    assert o.coefs.shape == g['coefs_head'].shape and np.array_equal(o.coefs, g['coefs_head'])
    assert np.all(o.coefs[:, 1, 0, :] == 0.0)              # m = 0 sine row
    o._repackage_cylindrical_coefficients()
    assert np.array_equal(o.C[1]['sin'][3], g['coefs_head'][:, 1, 1, 3])
    d = o._repackage_cylindrical_coefficients_compatibility()
    assert len(d) == len(o.T) and np.array_equal(d[o.T[2]].cos, g['coefs_head'][2, 0])


@pytest.mark.skipif(not os.path.exists(FIXTURE), reason='reference tree not mounted')
def test_reader_matches_reference_on_full_fixture():
    g = np.load(os.path.join(GOLDEN, 'outcoef_star.npz'))
    o = outcoef.OutCoef(FIXTURE, verbose=0)
    assert np.array_equal(o.T, g['T_full'])
    assert tuple(o.coefs.shape) == tuple(g['shape_full'])
    assert hashlib.sha256(np.ascontiguousarray(o.coefs).tobytes()).hexdigest() == str(g['sha_full'])
    # and live against the unmodified reference reader
    from oracle import refshim
    refshim.load()
    with contextlib.redirect_stdout(io.StringIO()):
        import exptool.io.outcoef as ref_outcoef
        r = ref_outcoef.OutCoef(FIXTURE)
    assert np.array_equal(r.T, o.T) and np.array_equal(r.coefs, o.coefs)


def test_writer_round_trip_cylinder_and_sphere(tmp_path):
    rng = np.random.default_rng(5)
    S, M, N, L = 7, 7, 18, 4
    T = np.cumsum(rng.random(S))
    cos = rng.standard_normal((S, M, N)); sin = rng.standard_normal((S, M, N)); sin[:, 0] = 0.0
    f = str(tmp_path / 'outcoef.star.gpu')
    outcoef.write_cylinder_outcoef(f, T[:4], cos[:4], sin[:4])
    outcoef.write_cylinder_outcoef(f, T[4:], cos[4:], sin[4:], append=True)
    o = outcoef.OutCoef(f, verbose=0)
    assert o.basis == 'Cylinder' and np.array_equal(o.T, T)
    assert np.array_equal(o.coefs[:, 0], cos) and np.array_equal(o.coefs[:, 1], sin)
    coef = rng.standard_normal((S, (L + 1) ** 2, N))
    f2 = str(tmp_path / 'outcoef.dark.gpu')
    outcoef.write_sphere_outcoef(f2, T, coef)
    o2 = outcoef.OutCoef(f2, verbose=0)
    assert o2.basis == 'SphereSL' and np.array_equal(o2.T, T) and np.array_equal(o2.coefs, coef)
    o2._repackage_spherical_coefficients()
    assert np.array_equal(o2.C[2][1]['sin'][5], coef[:, 4 + 2, 5])
    if os.path.exists(FIXTURE):                           # the reference reader accepts the written files
        from oracle import refshim
        refshim.load()
        with contextlib.redirect_stdout(io.StringIO()):
            import exptool.io.outcoef as ref_outcoef
            r, r2 = ref_outcoef.OutCoef(f), ref_outcoef.OutCoef(f2)
        assert np.array_equal(r.T, T) and np.array_equal(r.coefs[:, 0], cos) and np.array_equal(r.coefs[:, 1], sin)
        assert np.array_equal(r2.T, T) and np.array_equal(r2.coefs, coef)


def test_old_formats(tmp_path):
    rng = np.random.default_rng(6)
    S, mmax, nmax, lmax = 5, 3, 6, 2
    T = np.arange(S) * 0.25
    rows = rng.standard_normal((S, 2 * mmax + 1, nmax))
    f = str(tmp_path / 'old_cyl')
    with open(f, 'wb') as fh:
        for t in range(S):
            np.array([T[t]], dtype='<f8').tofile(fh); np.array([mmax, nmax], dtype='<u4').tofile(fh)
            rows[t].astype('<f8').tofile(fh)
    o = outcoef.OutCoef(f, verbose=0)
    assert o.basis == 'Cylinder' and np.array_equal(o.T, T)
    assert np.array_equal(o.coefs[:, 0, 0], rows[:, 0]) and np.array_equal(o.coefs[:, 0, 1:], rows[:, 1::2])
    assert np.array_equal(o.coefs[:, 1, 1:], rows[:, 2::2])
    nl = lmax * (lmax + 2) + 1
    srows = rng.standard_normal((S, nmax, nl))
    f2 = str(tmp_path / 'old_sph')
    with open(f2, 'wb') as fh:
        for t in range(S):
            np.array([b'Sphere SL basis'], dtype='S64').tofile(fh)
            np.array([T[t], 0.0667], dtype='<f8').tofile(fh); np.array([nmax, lmax], dtype='<u4').tofile(fh)
            srows[t].astype('<f8').tofile(fh)
    o2 = outcoef.OutCoef(f2, verbose=0)
    assert o2.basis == 'SphereSL' and np.array_equal(o2.T, T)
    assert np.array_equal(o2.coefs, np.transpose(srows, (0, 2, 1)))
    if os.path.exists(FIXTURE):
        from oracle import refshim
        refshim.load()
        with contextlib.redirect_stdout(io.StringIO()):
            import exptool.io.outcoef as ref_outcoef
            r, r2 = ref_outcoef.OutCoef(f), ref_outcoef.OutCoef(f2)
        assert np.array_equal(r.coefs, o.coefs) and np.array_equal(r2.coefs, o2.coefs)
        assert np.array_equal(r.T, o.T) and np.array_equal(r2.T, o2.T)
