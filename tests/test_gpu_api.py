"""
GPU tests through the reference-facing Python API (exptool_b200.basis.{eof,spheresl,potential},
exptool_b200.utils.integrate): same call shapes as the reference, compared with the golden
vectors the unmodified reference produced.  Run with -m gpu on the B200.
"""
import os
import tempfile

import numpy as np
import pytest

from helpers import load_golden, relerr, S

pytestmark = pytest.mark.gpu

TOL = 1e-10
ORBIT_TOL = 1e-8


@pytest.fixture(scope='module')
def api():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU tests selected but CUDA is not available')
    from exptool_b200.basis import eof, spheresl, potential
    from exptool_b200.utils import integrate, halo_methods
    from exptool_b200.io import particle
    return dict(eof=eof, spheresl=spheresl, potential=potential, integrate=integrate, halo_methods=halo_methods,
                particle=particle)


def _eof_file(tmp, meta):
    pe, T = S.make_eof_tables(meta['eof_params'], kind=meta['kind'], seed=meta['seed'])
    return S.write_eof_cache(os.path.join(tmp, 'eof.cache'), pe, T)


def _sl_files(tmp, meta, seed_offset=0):
    ps, ev, ef = S.make_sl_tables(meta['sl_params'], kind=meta['kind'], seed=meta['seed'] + seed_offset)
    sf = S.write_sl_cache(os.path.join(tmp, 'sl.cache'), ps, ev, ef)
    mf = S.write_hernquist_model(os.path.join(tmp, 'sl.model'), a=ps['scale'])
    return sf, mf


@pytest.mark.parametrize('name', ['eof_small_random_cmap1', 'eof_small_random_cmap0', 'eof_std_smooth'])
def test_eof_api(api, name):
    eof, particle = api['eof'], api['particle']
    d, meta = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        f = _eof_file(tmp, meta)
        potC, rfC, zfC, dC, potS, rfS, zfS, dS = eof.parse_eof(f)
        rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, dens = eof.eof_params(f)
        XMIN, XMAX, dX, YMIN, YMAX, dY = eof.set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ascale, HSCALE=hscale,
                                                              NUMX=numx, NUMY=numy, CMAP=cmap)
        P = S.ParticleSet(d['x'], d['y'], d['z'], d['m'])
        # eof.accumulate, both container branches (eof.py:525 / 582)
        c, s = eof.accumulate(P, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap)
        assert isinstance(c, np.ndarray) and c.shape == (mmax + 1, norder) and c.dtype == np.float64
        assert relerr(c, d['cos']) < TOL and relerr(s, d['sin']) < TOL
        H = particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = d['x'], d['y'], d['z'], d['m']
        c2, s2 = eof.accumulate(H, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap,
                                no_odd=True)          # no_odd has no effect in the reference either
        assert relerr(c2, d['cos']) < TOL
        # eof.make_coefficients_multi / eof.compute_coefficients
        c3, s3 = eof.make_coefficients_multi(P, 3, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale,
                                             hscale, cmap)
        assert relerr(c3, d['cos_multi3']) < TOL and relerr(s3, d['sin_multi3']) < TOL
        EO = eof.compute_coefficients(P, f, verbose=0)
        assert EO.mmax == mmax and EO.nmax == norder and EO.nbodies == d['x'].size
        assert relerr(EO.cos, d['cos']) < TOL and relerr(EO.sin, d['sin']) < TOL
        # NaN positions are moved to the origin (the reference's intent at eof.py:1190-1204)
        xn = d['x'].copy(); xn[10] = np.nan
        x0 = d['x'].copy(); y0 = d['y'].copy(); z0 = d['z'].copy(); x0[10] = y0[10] = z0[10] = 0.0
        En = eof.compute_coefficients(S.ParticleSet(xn, d['y'], d['z'], d['m']), f, verbose=0)
        E0 = eof.compute_coefficients(S.ParticleSet(x0, y0, z0, d['m']), f, verbose=0)
        assert relerr(En.cos, E0.cos) < 1e-13
        # eof.accumulated_eval_particles / compute_forces
        nf = meta['nforce']
        Pf = S.ParticleSet(d['x'][:nf], d['y'][:nf], d['z'][:nf], d['m'][:nf])
        kw = dict(potC=potC, rforceC=rfC, zforceC=zfC, potS=potS, rforceS=rfS, zforceS=zfS, rmin=XMIN, dR=dX,
                  zmin=YMIN, dZ=dY, numx=numx, numy=numy, MMAX=mmax, NMAX=norder, ASCALE=ascale, HSCALE=hscale,
                  CMAP=cmap, verbose=0)
        full = eof.accumulated_eval_particles(Pf, d['cos'], d['sin'], **kw)
        win = eof.accumulated_eval_particles(Pf, d['cos'], d['sin'], m1=1, m2=2, **kw)
        byfile = eof.accumulated_eval_particles(Pf, d['cos'], d['sin'], eof_file=f, verbose=0)
        EO.cos, EO.sin = d['cos'], d['sin']
        cf = eof.compute_forces(Pf, EO, verbose=0)
        for i in range(6):
            assert relerr(full[i], d['full'][i]) < TOL, i
            assert relerr(win[i], d['win12'][i]) < TOL, i
            assert relerr(byfile[i], d['full'][i]) < TOL, i
            assert relerr(cf[i], d['full'][i]) < TOL, i
        # eof.force_eval, scalar calls as the reference makes them, and batched
        variants = dict(full=dict(MMAX=mmax, NMAX=norder), trunc=dict(MMAX=max(mmax - 1, 1), NMAX=max(norder - 1, 1)),
                        noodd=dict(MMAX=mmax, NMAX=norder, no_odd=True), perturb=dict(MMAX=mmax, NMAX=norder, perturb=True))
        geo = dict(rmin=XMIN, dR=dX, zmin=YMIN, dZ=dY, numx=numx, numy=numy, ASCALE=ascale, HSCALE=hscale, CMAP=cmap)
        for vname, v in variants.items():
            ref = d['fe_' + vname]
            out = eof.force_eval(d['pt_r'], d['pt_z'], d['pt_phi'], d['cos'], d['sin'], potC, rfC, zfC, potS, rfS, zfS,
                                 **geo, **v)
            assert len(out) == ref.shape[1]
            for i in range(ref.shape[1]):
                assert relerr(out[i], ref[:, i]) < TOL, (vname, i)
            one = eof.force_eval(float(d['pt_r'][3]), float(d['pt_z'][3]), float(d['pt_phi'][3]), d['cos'], d['sin'],
                                 potC, rfC, zfC, potS, rfS, zfS, **geo, **v)
            assert all(np.ndim(o) == 0 for o in one)
            assert max(abs(float(one[i]) - ref[3, i]) for i in range(ref.shape[1])) <= TOL * np.max(np.abs(ref))
        if name == 'eof_small_random_cmap1':
            # VAR: jackknife partitions drawn with NumPy's global generator exactly as the reference draws them
            v = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'eof_var.npz'))
            vm = __import__('json').loads(str(v['meta']))
            for cont, key in ((H, 'holder'), (P, 'data')):
                np.random.seed(vm['seed'])
                cv, sv, cv2, sv2 = eof.accumulate(cont, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale,
                                                  hscale, cmap, VAR=vm['nvar'])
                assert cv2.shape == (vm['nvar'], mmax + 1, norder)
                assert relerr(cv, v['cos']) < TOL and relerr(sv, v['sin']) < TOL
                assert relerr(cv2, v['cos2_' + key]) < TOL and relerr(sv2, v['sin2_' + key]) < TOL
            np.random.seed(vm['seed'])
            EV = eof.compute_coefficients(P, f, verbose=0, VAR=vm['nvar'])
            assert relerr(EV.cos2, v['cos2_data']) < TOL


@pytest.mark.parametrize('name', ['sl_small_random_cmap1', 'sl_small_random_cmap0', 'sl_std_l4', 'sl_std_l6'])
def test_spheresl_api(api, name):
    spheresl, halo_methods, particle = api['spheresl'], api['halo_methods'], api['particle']
    d, meta = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        sf, mf = _sl_files(tmp, meta)
        H = particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = d['x'], d['y'], d['z'], d['m']
        c = spheresl.compute_coefficients_solitary(H, sf, mf, verbose=0)
        assert c.shape == d['coef'].shape and relerr(c, d['coef']) < TOL
        c2 = spheresl.compute_coefficients_solitary(H, sf, mf, verbose=0, no_odd=True)
        assert relerr(c2, d['coef_noodd']) < TOL
        SO = spheresl.compute_coefficients(S.ParticleSet(d['x'], d['y'], d['z'], d['m']), sf, mf, verbose=0)
        assert SO.lmax == meta['lmax'] and SO.nmax == meta['nmax'] and relerr(SO.expcoef, d['coef']) < TOL
        assert SO.expcoef.dtype == np.float64
        nf = meta['nforce']
        Pf = S.ParticleSet(d['x'][1:nf + 1], d['y'][1:nf + 1], d['z'][1:nf + 1], d['m'][1:nf + 1])
        for key, kw in (('allp', {}), ('allp_win12', dict(L1=1, L2=2)), ('allp_noodd', dict(NO_ODD=True))):
            out = spheresl.all_eval_particles(Pf, d['coef'], sf, mf, 0, **kw)
            assert len(out) == 8
            for j in range(8):            # incl. den0, den1 with the function's own quirks (App. C #8)
                assert relerr(out[j], d[key][j]) < TOL, (key, j)
        lmax, nmax, numr, cmap, rmin, rmax, scale, ltable, evtable, eftable = halo_methods.read_cached_table(sf)
        xi, rarr, p0, d0 = halo_methods.init_table(mf, numr, rmin, rmax, cmap=cmap, scale=scale)
        a = (d['pt_r'], d['pt_costh'], d['pt_phi'], d['coef'], xi, p0, d0, cmap, scale)
        for key, (l, n, no_odd) in dict(fe_full=(lmax, nmax, False), fe_trunc=(max(lmax - 1, 1), max(nmax - 2, 1), False),
                                        fe_noodd=(lmax, nmax, True)).items():
            out = spheresl.force_eval(*a, l, n, evtable, eftable, no_odd=no_odd)
            for i in range(5):
                assert relerr(out[i], d[key][:, i]) < TOL, (key, i)
        one = spheresl.force_eval(float(d['pt_r'][2]), float(d['pt_costh'][2]), float(d['pt_phi'][2]), d['coef'], xi, p0,
                                  d0, cmap, scale, lmax, nmax, evtable, eftable)
        assert all(np.ndim(o) == 0 for o in one)
        out = spheresl.all_eval(*a, lmax, nmax, evtable, eftable)
        for j in range(7):                # den0, den1 (total density), pot0, pot1, potr, pott, potp
            assert relerr(out[j], d['ae_full'][:, j]) < TOL, j


@pytest.mark.parametrize('name', ['field_small', 'field_std'])
def test_fields_and_leapfrog_api(api, name):
    potential, integrate = api['potential'], api['integrate']
    d, meta = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        ef = _eof_file(tmp, meta)
        sf, mf = _sl_files(tmp, meta, seed_offset=1)
        F = potential.make_fields(ef, sf, mf, d['cos'], d['sin'], d['coef'], halofac=meta['halofac'])
        out = F.return_forces_cart(d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
        cyl = F.return_forces_cyl(d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
        for i in range(8):
            assert relerr(out[i], d['cart_full'][:, i]) < TOL, i
            assert relerr(cyl[i], d['cyl_full'][:, i]) < TOL, i
        one = F.return_forces_cart(float(d['px'][5]), float(d['py'][5]), float(d['pz'][5]), rotpos=meta['rot_full'])
        assert len(one) == 8 and all(np.ndim(o) == 0 for o in one)
        F.set_field_parameters(no_odd=True, halo_l=2, halo_n=3, disk_m=2, disk_n=3)
        out = F.return_forces_cart(d['px'], d['py'], d['pz'], rotpos=meta['rot_trunc'])
        for i in range(8):
            assert relerr(out[i], d['cart_trunc'][:, i]) < TOL, i
        F.reset_field_parameters()
        # integrate.leapfrog_integrate, one orbit at a time as the reference is called
        nint, dt = meta['nint'], meta['dt']
        keys = ('X', 'Y', 'Z', 'VX', 'VY', 'VZ', 'P', 'FX', 'FY', 'FZ', 'TX', 'TY', 'VTX', 'VTY', 'T')
        for k in range(d['orbits'].shape[0]):
            Orb = integrate.leapfrog_integrate(F, nint, dt, d['pos0'][:, k], d['vel0'][:, k], rotfreq=meta['rotfreq'],
                                               force=True)
            for j, key in enumerate(keys):
                assert Orb[key].shape == (nint,)
                assert relerr(Orb[key], d['orbits'][k, j]) < ORBIT_TOL, (k, key)
        Ot = integrate.leapfrog_integrate(F, nint, dt, d['pos0'][:, 0], d['vel0'][:, 0], rotfreq=3.0, no_odd=True,
                                          halo_l=2, halo_n=4, disk_m=4, disk_n=5)
        assert 'FX' not in Ot
        for j, key in enumerate(('X', 'Y', 'Z', 'VX', 'VY', 'VZ', 'P', 'TX', 'TY', 'VTX', 'VTY', 'T')):
            assert relerr(Ot[key], d['orbit_trunc'][j]) < ORBIT_TOL, key
        # batched form == the per-orbit calls; apse counting stops orbits early
        F.reset_field_parameters()
        B = integrate.leapfrog_integrate_batch(F, nint, dt, d['pos0'], d['vel0'], rotfreq=meta['rotfreq'])
        for k in range(d['orbits'].shape[0]):
            assert abs(B['X'][k] - d['orbits'][k, 0, -1]) <= ORBIT_TOL * np.max(np.abs(d['orbits'][k, 0]))
        Oa = integrate.leapfrog_integrate(F, 4 * nint, dt, d['pos0'][:, 0], d['vel0'][:, 0], rotfreq=0.0, apse=True, ap_max=1)
        X, Y = Oa['X'], Oa['Y']
        r2 = X * X + Y * Y
        n = len(X)
        if n < 4 * nint:          # an apocentre was found: it is the step before the last one returned
            assert r2[n - 2] > r2[n - 3] and r2[n - 2] > r2[n - 1]


def test_fields_total_coefficients_in_memory(api):
    potential = api['potential']
    with tempfile.TemporaryDirectory() as tmp:
        ef, sf, mf = S.write_fixture_files(tmp, eof_params=dict(mmax=2, numx=16, numy=12, nmax=8, norder=3),
                                           sl_params=dict(lmax=2, nmax=4, numr=100))
        F = potential.Fields('memory', ef, sf, mf, verbose=0)
        with pytest.raises(IOError):
            F.total_coefficients()                         # no such PSP file (the reference fails the same way)
        disc = S.ParticleSet(*S.exponential_disc(3000, 1))
        halo = S.ParticleSet(*S.hernquist_halo(2000, 2))
        F.total_coefficients(disc=disc, halo=halo, halofac=2.0)
        F.prep_tables()
        a = F.return_forces_cart(0.01, 0.002, 0.0005)
        assert len(a) == 8 and all(np.isfinite(a))
        assert F.EOF.cos.shape == (3, 3) and F.SL.expcoef.shape == (9, 4) and F.halofac == 2.0


@pytest.mark.parametrize('name', ['field_small', 'field_std'])
def test_frozen_field_timestep_and_grids(api, name):
    """restore_field -> forces; compute_timestep; integrate_grid* against the reference's own outputs."""
    potential, integrate = api['potential'], api['integrate']
    d, meta = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        ef = _eof_file(tmp, meta)
        sf, mf = _sl_files(tmp, meta, seed_offset=1)
        F = potential.make_fields(ef, sf, mf, d['cos'], d['sin'], d['coef'], halofac=meta['halofac'])
        ff = os.path.join(tmp, 'frozen.field')
        F.save_field(ff)
        R = potential.restore_field(ff)
        R.set_field_parameters()
        out = R.return_forces_cart(d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
        for i in range(8):
            assert relerr(out[i], d['cart_restored'][:, i]) < TOL, i
        # the float32 geometry makes the restored field differ from the original at ~1e-6 (App. C #15)
        assert 1e-9 < relerr(out[0], d['cart_full'][:, 0]) < 1e-3
        # compute_timestep: scalar calls and batched.  (The golden values were taken while the reference Fields
        # still held the truncation of the preceding leapfrog call -- leapfrog_integrate sets the field
        # parameters and never resets them, integrate.py:89.)
        F.set_field_parameters(no_odd=True, halo_l=2, halo_n=4, disk_m=4, disk_n=5)
        for k in range(d['pos0'].shape[1]):
            dt = integrate.compute_timestep(F, d['pos0'][:, k], d['vel0'][:, k])
            assert abs(dt - d['timestep'][k]) <= 1e-10 * d['timestep'][k]
        dts = integrate.compute_timestep(F, d['pos0'], d['vel0'])
        assert relerr(dts, d['timestep']) < 1e-10
        F.reset_field_parameters()
        # orbit grids
        rads, vels = d['grid_rads'], d['grid_vels']
        g = integrate.integrate_grid(rads, vels, F, 40, 1.0e-5, meta['rotfreq'], False, -1, -1, 100., 1000, 0)
        assert g.shape == d['grid'].shape
        for k in range(7):
            assert relerr(g[:, :, k], d['grid'][:, :, k]) < ORBIT_TOL, k
        gy = integrate.integrate_grid_launchy(rads[:2], vels[:1], F, 30, 1.0e-5, 3.0, False, -1, -1, 50., 1000, 0)
        for k in range(7):
            assert relerr(gy[:, :, k], d['grid_y'][:, :, k]) < ORBIT_TOL, k
        g3 = integrate.integrate_grid_3D(rads[:2], vels, F, 24, 1.0e-5, meta['rotfreq'], False, -1, -1, 100., 1000,
                                         d['grid_zs'], d['grid_vzs'], 0)
        assert g3.shape == d['grid3'].shape
        for k in range(9):
            assert relerr(g3[..., k, :], d['grid3'][..., k, :]) < ORBIT_TOL, k
        # do_integrate_multi (integrate.py:290-333): the same grids gathered as one float32 array
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            gm = integrate.do_integrate_multi(rads, vels, F, 40, 1.0e-5, meta['rotfreq'], False, -1, -1, 100., 1000, nprocs=3)
            gm3 = integrate.do_integrate_multi(rads[:2], vels, F, 24, 1.0e-5, meta['rotfreq'], False, -1, -1, 100., 1000,
                                               threedee=True, zs=d['grid_zs'], vzs=d['grid_vzs'])
            gmy = integrate.do_integrate_multi(rads[:2], vels[:1], F, 30, 1.0e-5, 3.0, False, -1, -1, 50., 1000, launch='y')
        assert gm.dtype == np.float32 and np.array_equal(gm, g.astype('f4'))
        assert gm3.dtype == np.float32 and np.array_equal(gm3, g3.astype('f4'))
        assert np.array_equal(gmy, gy.astype('f4'))
        parts = integrate.redistribute_arrays(np.arange(11), 3)
        assert [len(p_) for p_ in parts] == [5, 3, 3] and np.array_equal(np.concatenate(parts), np.arange(11))
        assert integrate.re_form_orbit_arrays([g[:1], g[1:]]).shape == g.shape
        # run_time_mod (integrate.py:638-755), the workflow driver of the notebooks, on the frozen-field file: restore ->
        # rotfreq = -|omegap / 2 pi| -> do_integrate_multi -> omap text file (2-D and 3-D launch grids)
        omega = 2. * np.pi * abs(meta['rotfreq'])
        of2, of3 = os.path.join(tmp, 'omap.txt'), os.path.join(tmp, 'omap3.txt')
        with contextlib.redirect_stdout(io.StringIO()):
            integrate.run_time_mod(tmp + '/', 'run', ef, sf, mf, 7, rads, vels, 40, 1.0e-5, False, -1, -1, 100., 1000, 0,
                                   omegap=omega, orbitfile=of2, field_file=ff, field_file_name=ff)
            integrate.run_time_mod(tmp + '/', 'run', ef, sf, mf, 7, rads[:2], vels, 24, 1.0e-5, False, -1, -1, 100., 1000, 0,
                                   omegap=omega, orbitfile=of3, field_file=ff, field_file_name=ff, threedee=True,
                                   zs=d['grid_zs'], vzs=d['grid_vzs'])
            want2 = integrate.do_integrate_multi(rads, vels, R, 40, 1.0e-5, -abs(meta['rotfreq']), False, -1, -1, 100., 1000)
        ref2 = os.path.join(tmp, 'ref2.txt')
        with open(ref2, 'w') as fh:
            integrate.print_orbit_array(fh, want2)
        assert open(of2).read() == open(ref2).read() and os.path.getsize(of2) > 1000
        assert os.path.getsize(of3) > 1000 and len(open(of3).readlines()) == 2 * len(vels) * len(d['grid_zs']) * len(d['grid_vzs'])
        # ap_max = 0: the reference's loop takes no step (integrate.py:126)
        O0 = integrate.leapfrog_integrate(F, 20, 1.0e-5, d['pos0'][:, 0], d['vel0'][:, 0], ap_max=0, apse=True)
        assert len(O0['T']) == 1 and O0['X'][0] == d['pos0'][0, 0]


def test_eof_host_pipeline_matches_device(api):
    """The chunked H2D | kernels | D2H pipeline behind the host-array API (ops.EOFTables.accumulate_host /
    force_host, used by eof.make_coefficients_multi and eof.accumulated_eval_particles for large sets) gives
    the results of the one-shot device path, for pinned tensors and for plain NumPy inputs, ragged last chunk."""
    import torch
    from exptool_b200 import ops
    eof = api['eof']
    d, meta = load_golden('eof_std_smooth')
    from helpers import eof_tables, eof_geo_args
    pe, T, g = eof_tables(meta)
    n = 3 * 131072 + 1234
    x, y, z, m = S.exponential_disc(n, 97)
    tabs = (T['potC'], T['potS'], g['mmax'], g['norder']) + eof_geo_args(g) + (g['ascale'], g['hscale'], g['cmap'])
    E = eof.device_tables(*tabs, rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
    cd, sd = E.accumulate(x, y, z, m)
    cd, sd = cd.cpu().numpy(), sd.cpu().numpy()
    c1, s1 = eof.make_coefficients_multi((x, y, z, m), 1, *tabs)
    pin = tuple(torch.from_numpy(a).pin_memory() for a in (x, y, z, m))
    c2, s2 = eof.make_coefficients_multi(pin, 1, *tabs)
    for c, s in ((c1, s1), (c2, s2)):
        assert isinstance(c, np.ndarray) and c.shape == cd.shape
        assert relerr(c, cd) < 1e-13 and relerr(s, sd) < 1e-13
    E.contract(cd, sd)
    ref = E.force(x, y, z).cpu().numpy()
    kw = dict(potC=T['potC'], rforceC=T['rforceC'], zforceC=T['zforceC'], potS=T['potS'], rforceS=T['rforceS'],
              zforceS=T['zforceS'], rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'],
              numy=g['numy'], MMAX=g['mmax'], NMAX=g['norder'], ASCALE=g['ascale'], HSCALE=g['hscale'],
              CMAP=g['cmap'], verbose=0)
    for P in ((x, y, z, m), pin):
        out = eof.accumulated_eval_particles(P, cd, sd, **kw)
        assert len(out) == 6
        for k in range(6):
            assert isinstance(out[k], np.ndarray) and out[k].shape == (n,)
            assert relerr(out[k], ref[k]) < 1e-13
    # one upload per snapshot (option host_reuse): the evaluation that follows an accumulation of the SAME host arrays
    # runs from the copy that accumulation uploaded -- across table handles (accumulate-only vs with force tables) --
    # and gives the bits of the path that uploads again; a changed set (in place, same pointers) is uploaded again.
    saved = ops.get_option('host_reuse')
    try:
        ops.set_option('host_reuse', 0)
        eof.make_coefficients_multi(pin, 1, *tabs)
        base = eof.accumulated_eval_particles(pin, cd, sd, **kw)
        assert ops.get_option('host_reused_last') == 0
        ops.set_option('host_reuse', 1)
        eof.make_coefficients_multi(pin, 1, *tabs)
        again = eof.accumulated_eval_particles(pin, cd, sd, **kw)
        assert ops.get_option('host_reused_last') == 1
        for k in range(6):            # bit for bit: the cell sort is stable (round 2), so the evaluation path per particle is fixed
            assert np.array_equal(again[k], base[k]), k
        other = eof.accumulated_eval_particles((x, y, z, m), cd, sd, **kw)      # other host arrays: not the kept set
        assert ops.get_option('host_reused_last') == 0
        pin[0].mul_(-1.0); pin[1].mul_(-1.0)                                      # in-place rotation by pi: same pointers, new content
        moved = eof.accumulated_eval_particles(pin, cd, sd, **kw)
        assert ops.get_option('host_reused_last') == 0
        refm = E.force(-x, -y, z).cpu().numpy()
        for k in range(6):
            assert relerr(moved[k], refm[k]) < 1e-13, k
        eof.make_coefficients_multi(pin, 1, *tabs)                                # a new accumulation uploads the moved set
        moved2 = eof.accumulated_eval_particles(pin, cd, sd, **kw)
        assert ops.get_option('host_reused_last') == 1
        for k in range(6):
            assert np.array_equal(moved2[k], moved[k]), k
    finally:
        ops.set_option('host_reuse', saved)


@pytest.mark.parametrize('name', ['eof_dens_random', 'eof_dens_smooth'])
def test_eof_density_api(api, name):
    """eof.accumulated_eval_particles(density=True, eof_file=...) against the reference's own output on a dens=1
    cache file: p0, p, d0, d, fr, fp, fz, R (eof.py:1136-1142)."""
    eof = api['eof']
    d, meta = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        f = _eof_file(tmp, meta)
        assert eof.eof_params(f)[9] == 1
        P = S.ParticleSet(d['x'], d['y'], d['z'], d['m'])
        full = eof.accumulated_eval_particles(P, d['cos'], d['sin'], density=True, eof_file=f, verbose=0)
        win = eof.accumulated_eval_particles(P, d['cos'], d['sin'], m1=1, m2=2, density=True, eof_file=f, verbose=0)
        assert len(full) == 8 and len(win) == 8
        for i in range(8):
            assert isinstance(full[i], np.ndarray) and full[i].shape == d['x'].shape
            assert relerr(full[i], d['full'][i]) < TOL, i
            assert relerr(win[i], d['win12'][i]) < TOL, i
        # without density tables the flag is dropped and the six standard outputs come back (eof.py:1048-1050)
        six = eof.accumulated_eval_particles(P, d['cos'], d['sin'], eof_file=f, verbose=0)
        assert len(six) == 6 and relerr(six[1], d['full'][1]) < TOL


def test_building_blocks_api(api):
    """a5 return_bins, a6 get_pot, a11 accumulated_eval, a14/a15 radial matrices, a16 Legendre tables: the
    reference's helper functions, evaluated on the device, against the unmodified reference (blocks_small.npz)."""
    eof, spheresl, halo_methods = api['eof'], api['spheresl'], api['halo_methods']
    d, meta = load_golden('blocks_small')
    g = meta['geo']
    with tempfile.TemporaryDirectory() as tmp:
        f = _eof_file(tmp, meta)
        potC, rforceC, zforceC, densC, potS, rforceS, zforceS, densS = eof.parse_eof(f)
        geo = dict(rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'], numy=g['numy'],
                   ASCALE=g['ascale'], HSCALE=g['hscale'], CMAP=g['cmap'])
        X, Y, ix, iy = eof.return_bins(d['r'], d['z'], **geo)
        assert relerr(X, d['X']) < TOL and relerr(Y, d['Y']) < TOL
        assert np.array_equal(ix, d['ix']) and np.array_equal(iy, d['iy'])
        one = eof.return_bins(float(d['r'][9]), float(d['z'][9]), **geo)
        assert all(np.ndim(o) == 0 for o in one)
        assert np.allclose([float(o) for o in one], d['bins_scalar'], rtol=1e-12, atol=0)
        Vc, Vs = eof.get_pot(d['r'], d['z'], potC, potS, fac=1.0, MMAX=g['mmax'], NMAX=g['norder'], **geo)
        assert Vc.shape == d['Vc'].shape and relerr(Vc, d['Vc']) < TOL and relerr(Vs, d['Vs']) < TOL
        a = (d['r'], d['z'], d['phi'], d['cosc'], d['sinc'], potC, rforceC, zforceC, densC, potS, rforceS, zforceS, densS)
        kw = dict(geo, MMAX=g['mmax'], NMAX=g['norder'])
        for key, no_odd in (('ae', False), ('ae_noodd', True)):
            out = eof.accumulated_eval(*a, no_odd=no_odd, **kw)
            assert len(out) == 7
            for j in range(7):
                assert relerr(out[j], d[key][:, j]) < TOL, (key, j)
        one = eof.accumulated_eval(float(d['r'][3]), float(d['z'][3]), float(d['phi'][3]), *a[3:], **kw)
        assert all(np.ndim(o) == 0 for o in one) and abs(float(one[1]) - d['ae'][3, 1]) <= TOL * np.max(np.abs(d['ae'][:, 1]))
        # the fan-out helpers keep the reference's partition and concatenate in block order
        P = S.ParticleSet(d['r'], d['z'], d['phi'], np.ones_like(d['r']))
        hold = eof.redistribute_particles(P, 3)
        assert [len(h.xpos) for h in hold] == [14, 13, 13] and np.array_equal(np.concatenate([h.xpos for h in hold]), d['r'])
        sf, mf = _sl_files(tmp, meta, seed_offset=1)
        lmax, nmax, numr, cmap, rmin, rmax, scale, ltable, evtable, eftable = halo_methods.read_cached_table(sf)
        xi, rarr, p0, d0 = halo_methods.init_table(mf, numr, rmin, rmax, cmap=cmap, scale=scale)
        dens, force, pot = spheresl.get_halo_dens_pot_force(d['rad'], lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale)
        for got, key in ((dens, 'dens'), (force, 'force'), (pot, 'pot')):
            assert relerr(np.moveaxis(got, 2, 0), d[key]) < TOL, key
        potm = spheresl.get_halo_pot_matrix(float(d['rad'][10]), lmax, nmax, evtable, eftable, xi, p0, cmap, scale)
        assert potm.shape == (lmax + 1, nmax) and relerr(potm, d['potm'][10]) < TOL
        densm = spheresl.get_halo_dens(d['rad'], lmax, nmax, evtable, eftable, xi, d0, cmap, scale)
        assert relerr(np.moveaxis(densm, 2, 0), d['densm']) < TOL
    L = meta['leg_lmax']
    P, dP = spheresl.dlegendre_R(L, d['cth'])
    assert relerr(np.moveaxis(P, 2, 0), d['P2']) < 1e-14 and relerr(np.moveaxis(dP, 2, 0), d['dP']) < TOL
    P1 = spheresl.legendre_R(L, float(d['cth'][8]))
    assert P1.shape == (L + 1, L + 1) and relerr(P1, d['P'][8]) < 1e-14


def test_fields_table_fp32_option(api):
    """potential.Fields(table_fp32=True): the FP32-table mode through the reference-facing API; the option is scoped
    to that instance's calls (a second, FP64 instance keeps full parity in between)."""
    potential, integrate = api['potential'], api['integrate']
    d, meta = load_golden('field_small')
    with tempfile.TemporaryDirectory() as tmp:
        ef = _eof_file(tmp, meta); sf, mf = _sl_files(tmp, meta, seed_offset=1)
        Fs = [potential.make_fields(ef, sf, mf, d['cos'], d['sin'], d['coef'], halofac=meta['halofac'], table_fp32=fp32)
              for fp32 in (True, False)]
        F32, F64 = Fs
        a = F32.return_forces_cart(d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
        b = F64.return_forces_cart(d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
        e32 = max(relerr(a[i], d['cart_full'][:, i]) for i in range(8))
        e64 = max(relerr(b[i], d['cart_full'][:, i]) for i in range(8))
        assert e64 < TOL and 1e-12 < e32 < 1e-5, (e32, e64)
        o32 = integrate.leapfrog_integrate_batch(F32, 50, meta['dt'], d['pos0'], d['vel0'], rotfreq=meta['rotfreq'])
        o64 = integrate.leapfrog_integrate_batch(F64, 50, meta['dt'], d['pos0'], d['vel0'], rotfreq=meta['rotfreq'])
        assert 0 < relerr(o32['X'], o64['X']) < 1e-4


def test_fanout_mirrors_and_eval_particles(api):
    """The fan-out helpers the reference drives its Pool with -- eof.redistribute_particles / multi_accumulate /
    find_forces_multi / mix_outputs (eof.py:1333-1590) -- and spheresl.eval_particles (spheresl.py:502-528), against the
    goldens of the functions they wrap: block partials sum to the whole-set coefficients, per-block outputs concatenate to
    the whole-set outputs, the m window passed to find_forces_multi is ignored (eof.py:1528)."""
    eof, spheresl = api['eof'], api['spheresl']
    d, meta = load_golden('eof_small_random_cmap1')
    with tempfile.TemporaryDirectory() as tmp:
        f = _eof_file(tmp, meta)
        potC, rfC, zfC, dC, potS, rfS, zfS, dS = eof.parse_eof(f)
        rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, dens = eof.eof_params(f)
        XMIN, XMAX, dX, YMIN, YMAX, dY = eof.set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ascale, HSCALE=hscale,
                                                              NUMX=numx, NUMY=numy, CMAP=cmap)
        P = S.ParticleSet(d['x'], d['y'], d['z'], d['m'])
        tabs = (potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap)
        holders = eof.redistribute_particles(P, 3)
        sizes = [len(h.xpos) for h in holders]
        assert sum(sizes) == d['x'].size and sizes[0] >= sizes[1] == sizes[2]          # block 0 takes the remainder
        parts = eof.multi_accumulate(holders, 3, *tabs)
        assert len(parts) == 3 and all(len(pc) == 2 for pc in parts)
        csum = np.sum(np.array([pc[0] for pc in parts]), axis=0)                           # eof.py:1440
        ssum = np.sum(np.array([pc[1] for pc in parts]), axis=0)
        assert relerr(csum, d['cos_multi3']) < TOL and relerr(ssum, d['sin_multi3']) < TOL
        nf = meta['nforce']
        Pf = S.ParticleSet(d['x'][:nf], d['y'][:nf], d['z'][:nf], d['m'][:nf])
        out = eof.find_forces_multi(Pf, 2, d['cos'], d['sin'], potC, rfC, zfC, potS, rfS, zfS, XMIN, dX, YMIN, dY, numx, numy,
                                    mmax, norder, ascale, hscale, cmap, m1=1, m2=2)
        for i in range(6):
            assert relerr(out[i], d['full'][i]) < TOL, i                                   # window ignored: the full sum
        kw = dict(potC=potC, rforceC=rfC, zforceC=zfC, potS=potS, rforceS=rfS, zforceS=zfS, rmin=XMIN, dR=dX, zmin=YMIN,
                  dZ=dY, numx=numx, numy=numy, MMAX=mmax, NMAX=norder, ASCALE=ascale, HSCALE=hscale, CMAP=cmap, verbose=0)
        blocks = [eof.accumulated_eval_particles(h, d['cos'], d['sin'], **kw) for h in eof.redistribute_particles(Pf, 3)]
        mixed = eof.mix_outputs(blocks)
        assert len(mixed) == 6
        for i in range(6):
            assert mixed[i].shape == (nf,) and relerr(mixed[i], d['full'][i]) < TOL, i
    d, meta = load_golden('sl_std_l4')
    with tempfile.TemporaryDirectory() as tmp:
        sf, mf = _sl_files(tmp, meta)
        nf = meta['nforce']
        Pf = S.ParticleSet(d['x'][1:nf + 1], d['y'][1:nf + 1], d['z'][1:nf + 1], d['m'][1:nf + 1])
        for key, kw in (('allp', {}), ('allp_win12', dict(l1=1, l2=2)), ('allp_noodd', dict(no_odd=True))):
            out = spheresl.eval_particles(Pf, d['coef'], sf, mf, nprocs=4, verbose=0, **kw)
            assert len(out) == 8
            for j in range(8):
                assert relerr(out[j], d[key][j]) < TOL, (key, j)


def test_torch_library_custom_ops(api):
    """The hot path as PyTorch custom ops (north_star; exptool_b200/torch_ops.py): torch.ops.exptool_b200.* called directly
    with device tensors + table handles, against the same goldens as the reference-facing API."""
    import torch
    from exptool_b200 import ops, torch_ops
    from helpers import eof_tables, sl_tables
    T = torch.ops.exptool_b200
    for name in torch_ops.OPS:
        assert hasattr(T, name), name
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()
    d, meta = load_golden('eof_std_smooth')
    pe, Tt, g = eof_tables(meta)
    E = ops.EOFTables(Tt['potC'], Tt['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'],
                      g['numy'], g['ascale'], g['hscale'], g['cmap'], rforceC=Tt['rforceC'], zforceC=Tt['zforceC'],
                      rforceS=Tt['rforceS'], zforceS=Tt['zforceS'])
    x, y, z, m = [dev(d[k]) for k in 'xyzm']
    cs = T.eof_accumulate(E.handle, x, y, z, m, g['mmax'], g['norder'])
    assert cs.shape == (2, g['mmax'] + 1, g['norder']) and cs.is_cuda
    assert relerr(cs[0].cpu().numpy(), d['cos']) < TOL and relerr(cs[1].cpu().numpy(), d['sin']) < TOL
    T.eof_contract(E.handle, dev(d['cos']), dev(d['sin']), 0, g['mmax'], g['norder'], False)
    nf = meta['nforce']
    out = T.eof_force(E.handle, x[:nf].contiguous(), y[:nf].contiguous(), z[:nf].contiguous()).cpu().numpy()
    for i in range(6):
        assert relerr(out[i], d['full'][i]) < TOL, i
    d, meta = load_golden('sl_std_l6')
    ps, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
    x, y, z, m = [dev(d[k]) for k in 'xyzm']
    coef = T.sl_accumulate(H.handle, x, y, z, m, H.nrow, H.nmax, False)
    assert relerr(coef.cpu().numpy(), d['coef']) < TOL
    T.sl_contract(H.handle, dev(d['coef']), 0, ps['lmax'], ps['nmax'], False)
    nf = meta['nforce']
    out = T.sl_force(H.handle, x[1:nf + 1].contiguous(), y[1:nf + 1].contiguous(), z[1:nf + 1].contiguous()).cpu().numpy()
    for j, k in enumerate((2, 3, 4, 5, 6, 7)):                # all_eval_particles: den0 den1 pot0 pot1 potr pott potp r
        assert relerr(out[j], d['allp'][k]) < TOL, j
    d, meta = load_golden('field_std')
    pe, Tt, g = eof_tables(meta)
    ps, ev, ef, xi, p0, d0 = sl_tables(meta, seed_offset=1)
    E = ops.EOFTables(Tt['potC'], Tt['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'],
                      g['numy'], g['ascale'], g['hscale'], g['cmap'], rforceC=Tt['rforceC'], zforceC=Tt['zforceC'],
                      rforceS=Tt['rforceS'], zforceS=Tt['zforceS'])
    H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
    T.eof_contract(E.handle, dev(d['cos']), dev(d['sin']), 0, g['mmax'], g['norder'], False)
    T.sl_contract(H.handle, dev(meta['halofac'] * d['coef']), 0, ps['lmax'], ps['nmax'], False)
    out = T.field_force_cart(E.handle, H.handle, dev(d['px']), dev(d['py']), dev(d['pz']), float(meta['rot_full'])).cpu().numpy()
    for i in range(8):
        assert relerr(out[i], d['cart_full'][:, i]) < TOL, i
    outc = T.field_force_cyl(E.handle, H.handle, dev(d['px']), dev(d['py']), dev(d['pz']), float(meta['rot_full']))
    assert outc.shape == out.shape
    st = T.leapfrog(E.handle, H.handle, torch.cat([dev(d['pos0']), dev(d['vel0'])]).contiguous(), int(meta['nint']),
                    float(meta['dt']), float(meta['rotfreq'])).cpu().numpy()
    for k in range(d['orbits'].shape[0]):
        for j in range(6):
            assert abs(st[j, k] - d['orbits'][k, j, -1]) <= ORBIT_TOL * np.max(np.abs(d['orbits'][k, j])), (k, j)
    with pytest.raises(Exception):                                # no CPU kernel behind the ops
        T.eof_force(E.handle, torch.zeros(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                    torch.zeros(4, dtype=torch.float64))


def test_sl_host_pipeline_matches_device(api):
    """spheresl.compute_coefficients through the chunked host pipeline (bfe_sl_accumulate_host) and all_eval_particles with
    one upload / one copy out, on a set large enough for several chunks, against the one-shot device path."""
    import torch
    from exptool_b200 import ops
    spheresl = api['spheresl']
    d, meta = load_golden('sl_std_l4')
    with tempfile.TemporaryDirectory() as tmp:
        sf, mf = _sl_files(tmp, meta)
        n = 2 * 131072 + 4321
        x, y, z, m = S.hernquist_halo(n, 23)
        H, T = spheresl.device_tables_from_files(sf, mf)
        want = H.accumulate(x, y, z, m).cpu().numpy()
        for P in (S.ParticleSet(x, y, z, m), tuple(torch.from_numpy(a).pin_memory() for a in (x, y, z, m))):
            SO = spheresl.compute_coefficients(P, sf, mf, verbose=0)
            assert SO.expcoef.shape == want.shape and relerr(SO.expcoef, want) < 1e-13
        got = H.accumulate_host(x, y, z, m, no_odd=True)
        assert relerr(got, H.accumulate(x, y, z, m, no_odd=True).cpu().numpy()) < 1e-13
        H.contract(want)
        ref = H.force(x, y, z).cpu().numpy()
        pin = tuple(torch.from_numpy(a).pin_memory() for a in (x, y, z))
        H.accumulate_host(pin[0], pin[1], pin[2], torch.from_numpy(m).pin_memory())
        for args in ((x, y, z), pin):
            out = H.force_host(*args)
            assert out.shape == (6, n)
            for k in range(6):
                assert relerr(out[k], ref[k]) < 1e-13, k
        assert ops.get_option('host_reused_last') == 1          # the pinned set is the one just accumulated
        full = spheresl.all_eval_particles(S.ParticleSet(x[:5000], y[:5000], z[:5000], m[:5000]), want, sf, mf, 0)
        assert len(full) == 8
        for k in range(6):
            assert relerr(full[2 + k], ref[k][:5000]) < 1e-13, k
