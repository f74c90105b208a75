"""
Pin the oracle (oracle/oracle_np.py) against the golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

from helpers import load_golden, relerr, eof_tables, sl_tables, eof_geo_args, build_field, O

TOL = 1e-12     # vectorisation changes FP64 summation order only

EOF_CASES = ['eof_small_random_cmap1', 'eof_small_random_cmap0', 'eof_std_smooth']
SL_CASES = ['sl_small_random_cmap1', 'sl_small_random_cmap0', 'sl_std_l4', 'sl_std_l6']
FIELD_CASES = ['field_small', 'field_std']


@pytest.mark.parametrize('name', EOF_CASES)
def test_eof_accumulate(name):
    d, meta = load_golden(name)
    p, T, g = eof_tables(meta)
    for k, v in meta['geo'].items():                 # geometry agrees with eof.set_table_params
        assert abs(g[k] - v) <= 1e-15 * max(1.0, abs(v)), k
    c, s = O.eof_accumulate(d['x'], d['y'], d['z'], d['m'], T['potC'], T['potS'], g['mmax'], g['norder'],
                            *eof_geo_args(g), g['ascale'], g['hscale'], g['cmap'])
    assert relerr(c, d['cos']) < TOL
    assert relerr(s, d['sin']) < TOL
    # the reference's own 3-process split differs from its 1-process run by summation order only
    assert relerr(d['cos_multi3'], d['cos']) < TOL


@pytest.mark.parametrize('name', EOF_CASES)
def test_eof_force_particles(name):
    d, meta = load_golden(name)
    p, T, g = eof_tables(meta)
    nf = meta['nforce']
    args = (d['x'][:nf], d['y'][:nf], d['z'][:nf], d['cos'], d['sin'], T['potC'], T['rforceC'], T['zforceC'],
            T['potS'], T['rforceS'], T['zforceS'], *eof_geo_args(g), g['mmax'], g['norder'],
            g['ascale'], g['hscale'], g['cmap'])
    full = O.eof_force_particles(*args)
    win = O.eof_force_particles(*args, m1=1, m2=2)
    for i in range(6):
        assert relerr(full[i], d['full'][i]) < TOL, i
        assert relerr(win[i], d['win12'][i]) < TOL, i


@pytest.mark.parametrize('name', EOF_CASES)
def test_eof_force_eval(name):
    d, meta = load_golden(name)
    p, T, g = eof_tables(meta)
    variants = dict(full=(g['mmax'], g['norder'], False, False),
                    trunc=(max(g['mmax'] - 1, 1), max(g['norder'] - 1, 1), False, False),
                    noodd=(g['mmax'], g['norder'], True, False),
                    perturb=(g['mmax'], g['norder'], False, True))
    for vname, (M, N, no_odd, perturb) in variants.items():
        out = O.eof_force_eval(d['pt_r'], d['pt_z'], d['pt_phi'], d['cos'], d['sin'], T['potC'], T['rforceC'],
                               T['zforceC'], T['potS'], T['rforceS'], T['zforceS'], *eof_geo_args(g), M, N,
                               g['ascale'], g['hscale'], g['cmap'], no_odd=no_odd, perturb=perturb)
        ref = d['fe_' + vname]
        for i in range(ref.shape[1]):
            assert relerr(out[i], ref[:, i]) < TOL, (vname, i)


@pytest.mark.parametrize('name', SL_CASES)
def test_sl_tables_and_accumulate(name):
    d, meta = load_golden(name)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    assert relerr(xi, d['xi']) < 1e-15
    assert relerr(p0, d['p0']) < 1e-14
    assert relerr(d0, d['d0']) < 1e-14
    c = O.sl_accumulate(d['x'], d['y'], d['z'], d['m'], p['lmax'], p['nmax'], ev, ef, xi, p0, p['cmap'], p['scale'])
    assert relerr(c, d['coef']) < TOL
    c2 = O.sl_accumulate(d['x'], d['y'], d['z'], d['m'], p['lmax'], p['nmax'], ev, ef, xi, p0, p['cmap'], p['scale'],
                         no_odd=True)
    assert relerr(c2, d['coef_noodd']) < TOL


@pytest.mark.parametrize('name', SL_CASES)
def test_sl_eval_particles(name):
    d, meta = load_golden(name)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    nf = meta['nforce']
    a = (d['x'][1:nf + 1], d['y'][1:nf + 1], d['z'][1:nf + 1], d['coef'], p['lmax'], p['nmax'], ev, ef, xi, p0, d0,
         p['cmap'], p['scale'])
    for key, kw in (('allp', {}), ('allp_win12', dict(L1=1, L2=2)), ('allp_noodd', dict(NO_ODD=True))):
        out = O.sl_all_eval_particles(*a, **kw)
        ref = d[key]          # den0,den1,pot0,pot1,potr,pott,potp,rr
        for i, j in enumerate((2, 3, 4, 5, 6, 7)):
            assert relerr(out[i], ref[j]) < TOL, (key, j)
        # density outputs with the function's own quirks (spheresl.py:1323 legs[1][m], 1351 densfac = pi/4)
        outd = O.sl_all_eval_particles(*a, density=True, **kw)
        for j in range(8):
            assert relerr(outd[j], ref[j]) < TOL, (key, 'density', j)


@pytest.mark.parametrize('name', SL_CASES)
def test_sl_force_eval(name):
    d, meta = load_golden(name)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    L, N = p['lmax'], p['nmax']
    a = (d['pt_r'], d['pt_costh'], d['pt_phi'], d['coef'], xi, p0, d0, p['cmap'], p['scale'])
    for key, (l, n, no_odd) in dict(fe_full=(L, N, False), fe_trunc=(max(L - 1, 1), max(N - 2, 1), False),
                                    fe_noodd=(L, N, True)).items():
        out = O.sl_force_eval(*a, l, n, ev, ef, no_odd=no_odd)
        for i in range(5):
            assert relerr(out[i], d[key][:, i]) < TOL, (key, i)
    out = O.sl_all_eval(*a, L, N, ev, ef)     # den0,den1,pot0,pot1,potr,pott,potp
    for i, j in enumerate((2, 3, 4, 5, 6)):
        assert relerr(out[i], d['ae_full'][:, j]) < TOL, j
    outd = O.sl_all_eval(*a, L, N, ev, ef, density=True)     # den1 includes the monopole (spheresl.py:1046)
    for j in range(7):
        assert relerr(outd[j], d['ae_full'][:, j]) < TOL, ('density', j)


@pytest.mark.parametrize('name', FIELD_CASES)
def test_fields_forces_cart(name):
    d, meta = load_golden(name)
    F = build_field(meta, d)
    out = O.fields_forces_cart(F, d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
    for i in range(8):
        assert relerr(out[i], d['cart_full'][:, i]) < TOL, i
    out = O.fields_forces_cyl(F, d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
    for i in range(8):
        assert relerr(out[i], d['cyl_full'][:, i]) < TOL, i
    F.set_field_parameters(no_odd=True, halo_l=2, halo_n=3, disk_m=2, disk_n=3)
    out = O.fields_forces_cart(F, d['px'], d['py'], d['pz'], rotpos=meta['rot_trunc'])
    for i in range(8):
        assert relerr(out[i], d['cart_trunc'][:, i]) < TOL, i


@pytest.mark.parametrize('name', FIELD_CASES)
def test_leapfrog(name):
    d, meta = load_golden(name)
    F = build_field(meta, d)
    nint = meta['nint']
    pos, vel, pot, traj = O.leapfrog(F, nint, meta['dt'], d['pos0'], d['vel0'], rotfreq=meta['rotfreq'],
                                     keep_trajectory=True)
    ref = d['orbits']        # (norb, 15, nint): X Y Z VX VY VZ P FX FY FZ TX TY VTX VTY T
    for k in range(ref.shape[0]):
        for j, key in enumerate(('X', 'Y', 'Z', 'VX', 'VY', 'VZ', 'P', 'FX', 'FY', 'FZ')):
            assert relerr(traj[key][:, k], ref[k, j]) < 1e-9, (k, key)
    F2 = build_field(meta, d)
    pos, vel, pot, traj = O.leapfrog(F2, nint, meta['dt'], d['pos0'][:, :1], d['vel0'][:, :1], rotfreq=3.0,
                                     no_odd=True, halo_l=2, halo_n=4, disk_m=4, disk_n=5, keep_trajectory=True)
    for j, key in enumerate(('X', 'Y', 'Z', 'VX', 'VY', 'VZ', 'P')):
        assert relerr(traj[key][:, 0], d['orbit_trunc'][j]) < 1e-9, key


@pytest.mark.parametrize('name', ['eof_dens_random', 'eof_dens_smooth'])
def test_eof_density_particles(name):
    """density=True outputs (eof.py:1106, 1122, 1136-1142) of the unmodified reference on a dens=1 cache file."""
    d, meta = load_golden(name)
    p, T, g = eof_tables(meta)
    assert p['dens'] == 1
    args = (d['x'], d['y'], d['z'], d['cos'], d['sin'], T['potC'], T['rforceC'], T['zforceC'],
            T['potS'], T['rforceS'], T['zforceS'], *eof_geo_args(g), g['mmax'], g['norder'],
            g['ascale'], g['hscale'], g['cmap'])
    full = O.eof_force_particles(*args, densC=T['densC'], densS=T['densS'])
    win = O.eof_force_particles(*args, m1=1, m2=2, densC=T['densC'], densS=T['densS'])
    assert len(full) == 8
    for i in range(8):
        assert relerr(full[i], d['full'][i]) < TOL, i
        assert relerr(win[i], d['win12'][i]) < TOL, i


# ---------------------------------------------------------------------------
# per-point building blocks (a5, a6, a11 accumulated_eval, a14-a16): tests/golden/blocks_small.npz
# ---------------------------------------------------------------------------
def _blocks():
    d, meta = load_golden('blocks_small')
    pe, T, g = eof_tables(meta)
    geo = (g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'])
    return d, meta, T, g, geo


def test_blocks_eof_bins_and_get_pot():
    d, meta, T, g, geo = _blocks()
    X, Y, ix, iy = O.eof_return_bins(d['r'].copy(), d['z'], *geo, g['ascale'], g['hscale'], g['cmap'])
    assert relerr(X, d['X']) < 1e-14 and relerr(Y, d['Y']) < 1e-14
    assert np.array_equal(ix, d['ix']) and np.array_equal(iy, d['iy'])
    Vc, Vs = O.eof_get_pot(d['r'], d['z'], T['potC'], T['potS'], *geo, g['ascale'], g['hscale'], g['cmap'])
    assert relerr(Vc, d['Vc']) < TOL and relerr(Vs, d['Vs']) < TOL


def test_blocks_eof_accumulated_eval():
    d, meta, T, g, geo = _blocks()
    a = (d['r'], d['z'], d['phi'], d['cosc'], d['sinc'], T['potC'], T['rforceC'], T['zforceC'], T['densC'], T['potS'],
         T['rforceS'], T['zforceS'], T['densS'], *geo, g['mmax'], g['norder'], g['ascale'], g['hscale'], g['cmap'])
    for key, no_odd in (('ae', False), ('ae_noodd', True)):
        out = O.eof_accumulated_eval(*a, no_odd=no_odd)
        for j in range(7):
            assert relerr(out[j], d[key][:, j]) < TOL, (key, j)


def test_blocks_sl_radial_and_legendre():
    d, meta = load_golden('blocks_small')
    p, ev, ef, xi, p0, d0 = sl_tables(meta, seed_offset=1)
    dens, force, pot = O.sl_dens_pot_force(d['rad'], p['lmax'], p['nmax'], ev, ef, xi, d0, p0, p['cmap'], p['scale'])
    for got, key in ((dens, 'dens'), (force, 'force'), (pot, 'pot'), (pot, 'potm'), (dens, 'densm')):
        assert relerr(np.moveaxis(got, 2, 0), d[key]) < TOL, key
    L = meta['leg_lmax']
    P, dP = O.dlegendre_R(L, d['cth'])
    assert relerr(np.moveaxis(P, 2, 0), d['P']) < 1e-14
    assert relerr(np.moveaxis(P, 2, 0), d['P2']) < 1e-14
    assert relerr(np.moveaxis(dP, 2, 0), d['dP']) < 1e-12
