"""
make_golden.py -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Each .npz holds the inputs (particle arrays, table-generator parameters as JSON)
and the outputs of the reference's own functions (file:line cited per case).
Table VALUES are not stored: they are regenerated bit-identically by
exptool_b200.synthetic from the stored parameters (analytic formulas or
numpy.random.default_rng(seed) streams), written to disk in the reference's cache
formats and read back through the reader under test.
"""
import io
import json
import os
import sys
import tempfile
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refshim                      # noqa: E402
from exptool_b200 import synthetic as S         # noqa: E402

R = refshim.load()
eof, spheresl, potential = R['eof'], R['spheresl'], R['potential']
halo_methods, integrate, particle = R['halo_methods'], R['integrate'], R['particle']

SMALL_EOF = dict(mmax=3, numx=12, numy=10, nmax=8, norder=4)
SMALL_SL = dict(lmax=3, nmax=5, numr=60)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def edge_particles(rng, n, rscale, zscale):
    """random particles plus the edge cases the reference's bin logic distinguishes."""
    x = rng.normal(0, rscale, n); y = rng.normal(0, rscale, n); z = rng.normal(0, zscale, n)
    m = rng.random(n) / n
    # origin, on-axis, far outside the table radially and vertically, negative/zero z
    x[0], y[0], z[0] = 0.0, 0.0, 0.0
    x[1], y[1], z[1] = 0.0, 0.0, 3 * zscale
    x[2], y[2], z[2] = 50 * rscale, 0.0, 0.0
    x[3], y[3], z[3] = 0.3 * rscale, -0.2 * rscale, 500 * zscale
    x[4], y[4], z[4] = -1e3 * rscale, 1e3 * rscale, -500 * zscale
    x[5], y[5], z[5] = 1e-9, -1e-9, -1e-12
    x[6], y[6], z[6] = -rscale, 0.0, 0.0           # phi = pi
    x[7], y[7], z[7] = 0.0, -rscale, 1e-9          # phi = -pi/2
    return x, y, z, m


def eof_setup(tmp, eparams, kind, seed):
    pe, T = S.make_eof_tables(eparams, kind=kind, seed=seed)
    f = S.write_eof_cache(os.path.join(tmp, 'eof.cache'), pe, T)
    tabs = quiet(eof.parse_eof, f)
    rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, dens = eof.eof_params(f)
    XMIN, XMAX, dX, YMIN, YMAX, dY = eof.set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ascale, HSCALE=hscale,
                                                          NUMX=numx, NUMY=numy, CMAP=cmap)
    geo = dict(XMIN=float(XMIN), dX=float(dX), YMIN=float(YMIN), dY=float(dY), numx=int(numx), numy=int(numy),
               mmax=int(mmax), norder=int(norder), ascale=float(ascale), hscale=float(hscale), cmap=int(cmap))
    return f, tabs, geo


def case_eof(name, eparams, kind, seed, npart, rscale, zscale):
    rng = np.random.default_rng(seed + 100)
    with tempfile.TemporaryDirectory() as tmp:
        f, tabs, g = eof_setup(tmp, eparams, kind, seed)
        potC, rfC, zfC, dC, potS, rfS, zfS, dS = tabs
        x, y, z, m = edge_particles(rng, npart, rscale, zscale)
        P = S.ParticleSet(x, y, z, m)
        # eof.accumulate, .data branch (eof.py:580-640)
        cosd, sind = eof.accumulate(P, potC, potS, g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                                    g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'])
        # eof.accumulate, holder branch (eof.py:525-577)
        H = particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = x, y, z, m
        cosh, sinh = eof.accumulate(H, potC, potS, g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                                    g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'], no_odd=True)
        assert np.array_equal(cosd, cosh) and np.array_equal(sind, sinh)
        # eof.make_coefficients_multi with 3 workers (eof.py:1415-1455)
        cos3, sin3 = quiet(eof.make_coefficients_multi, P, 3, potC, potS, g['mmax'], g['norder'], g['XMIN'], g['dX'],
                           g['YMIN'], g['dY'], g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'])
        # eof.accumulated_eval_particles (eof.py:989-1144), full window and m1..m2 window
        kw = dict(potC=potC, rforceC=rfC, zforceC=zfC, potS=potS, rforceS=rfS, zforceS=zfS,
                  rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'], numy=g['numy'],
                  MMAX=g['mmax'], NMAX=g['norder'], ASCALE=g['ascale'], HSCALE=g['hscale'], CMAP=g['cmap'], verbose=0)
        nf = min(npart, 160)
        Pf = S.ParticleSet(x[:nf], y[:nf], z[:nf], m[:nf])
        full = eof.accumulated_eval_particles(Pf, cosd, sind, **kw)
        win = eof.accumulated_eval_particles(Pf, cosd, sind, m1=1, m2=2, **kw)
        # eof.force_eval at scalar points (eof.py:756-870): full, truncated, no_odd, perturb
        npt = min(npart, 64)
        rr = np.sqrt(x[:npt] ** 2 + y[:npt] ** 2) + 1e-15
        ph = np.arctan2(y[:npt], x[:npt]) + 0.37
        fe = {}
        variants = dict(full=dict(MMAX=g['mmax'], NMAX=g['norder'], no_odd=False, perturb=False),
                        trunc=dict(MMAX=max(g['mmax'] - 1, 1), NMAX=max(g['norder'] - 1, 1), no_odd=False, perturb=False),
                        noodd=dict(MMAX=g['mmax'], NMAX=g['norder'], no_odd=True, perturb=False),
                        perturb=dict(MMAX=g['mmax'], NMAX=g['norder'], no_odd=False, perturb=True))
        for vname, v in variants.items():
            rows = []
            for i in range(npt):
                out = eof.force_eval(rr[i], z[i], ph[i], cosd, sind, potC, rfC, zfC, potS, rfS, zfS,
                                     rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'], numy=g['numy'],
                                     ASCALE=g['ascale'], HSCALE=g['hscale'], CMAP=g['cmap'], **v)
                rows.append([float(o) for o in out])
            fe[vname] = np.array(rows)
    np.savez_compressed(os.path.join(HERE, name + '.npz'),
                        meta=json.dumps(dict(eof_params=eparams, kind=kind, seed=seed, geo=g, nforce=nf, npoint=npt)),
                        x=x, y=y, z=z, m=m, cos=cosd, sin=sind, cos_multi3=cos3, sin_multi3=sin3,
                        full=np.array(full), win12=np.array(win), pt_r=rr, pt_z=z[:npt], pt_phi=ph,
                        **{'fe_' + k: v for k, v in fe.items()})
    print('wrote', name)


def sl_setup(tmp, sparams, kind, seed):
    ps, ev, ef = S.make_sl_tables(sparams, kind=kind, seed=seed)
    sf = S.write_sl_cache(os.path.join(tmp, 'sl.cache'), ps, ev, ef)
    mf = S.write_hernquist_model(os.path.join(tmp, 'sl.model'), a=ps['scale'])
    return sf, mf, ps


def case_sl(name, sparams, kind, seed, npart):
    rng = np.random.default_rng(seed + 200)
    with tempfile.TemporaryDirectory() as tmp:
        sf, mf, ps = sl_setup(tmp, sparams, kind, seed)
        lmax, nmax, numr, cmap, rmin, rmax, scale, ltable, evtable, eftable = halo_methods.read_cached_table(sf)
        xi, rarr, p0, d0 = quiet(halo_methods.init_table, mf, numr, rmin, rmax, cmap=cmap, scale=scale)
        x, y, z, m = S.hernquist_halo(npart, seed + 300, a=ps['scale'], rmax=ps['rmax'])
        m = rng.random(npart) / npart
        # edge cases: origin (r floor 1e-10), on the z axis (costh=+-1), beyond rmax, tiny r
        x[0], y[0], z[0] = 0., 0., 0.
        x[1], y[1], z[1] = 0., 0., 0.2
        x[2], y[2], z[2] = 0., 0., -0.1
        x[3], y[3], z[3] = 5.0, -7.0, 3.0
        x[4], y[4], z[4] = 1e-7, 1e-7, -1e-7
        x[5], y[5], z[5] = -0.05, 0.0, 0.0
        H = particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = x, y, z, m
        # spheresl.compute_coefficients_solitary (spheresl.py:567-656)
        coef = quiet(spheresl.compute_coefficients_solitary, H, sf, mf, verbose=0, no_odd=False)
        coef_noodd = quiet(spheresl.compute_coefficients_solitary, H, sf, mf, verbose=0, no_odd=True)
        # spheresl.all_eval_particles (spheresl.py:1240-1362); skip the origin (costh = 0/0 -> NaN)
        nf = min(npart - 1, 120)
        Pf = S.ParticleSet(x[1:nf + 1], y[1:nf + 1], z[1:nf + 1], m[1:nf + 1])
        allp = quiet(spheresl.all_eval_particles, Pf, coef, sf, mf, 0)
        allp_win = quiet(spheresl.all_eval_particles, Pf, coef, sf, mf, 0, L1=1, L2=2, NO_ODD=False)
        allp_noodd = quiet(spheresl.all_eval_particles, Pf, coef, sf, mf, 0, NO_ODD=True)
        # spheresl.force_eval / all_eval at points (spheresl.py:1107-1234 / 987-1102)
        npt = min(npart - 1, 64)
        r3 = np.sqrt(x[1:npt + 1] ** 2 + y[1:npt + 1] ** 2 + z[1:npt + 1] ** 2) + 1e-15
        cth = z[1:npt + 1] / r3
        ph = np.arctan2(y[1:npt + 1], x[1:npt + 1]) - 0.21
        fe_full, fe_trunc, fe_noodd, ae_full = [], [], [], []
        for i in range(npt):
            a = (r3[i], cth[i], ph[i], coef, xi, p0, d0, cmap, scale)
            fe_full.append(spheresl.force_eval(*a, lmax, nmax, evtable, eftable))
            fe_trunc.append(spheresl.force_eval(*a, max(lmax - 1, 1), max(nmax - 2, 1), evtable, eftable))
            fe_noodd.append(spheresl.force_eval(*a, lmax, nmax, evtable, eftable, no_odd=True))
            ae_full.append(spheresl.all_eval(*a, lmax, nmax, evtable, eftable))
    np.savez_compressed(os.path.join(HERE, name + '.npz'),
                        meta=json.dumps(dict(sl_params=sparams, kind=kind, seed=seed, nforce=nf, npoint=npt,
                                             lmax=int(lmax), nmax=int(nmax))),
                        x=x, y=y, z=z, m=m, xi=xi, p0=p0, d0=d0, coef=np.asarray(coef, dtype=np.float64),
                        coef_noodd=np.asarray(coef_noodd, dtype=np.float64),
                        allp=np.array(allp), allp_win12=np.array(allp_win), allp_noodd=np.array(allp_noodd),
                        pt_r=r3, pt_costh=cth, pt_phi=ph, fe_full=np.array(fe_full, dtype=np.float64),
                        fe_trunc=np.array(fe_trunc, dtype=np.float64), fe_noodd=np.array(fe_noodd, dtype=np.float64),
                        ae_full=np.array(ae_full, dtype=np.float64))
    print('wrote', name)


def case_field(name, eparams, sparams, kind, seed, ndisc, nhalo, npts, norb, nint):
    """Fields.return_forces_cart (potential.py:445-497) and integrate.leapfrog_integrate (integrate.py:53-190)."""
    rng = np.random.default_rng(seed + 400)
    with tempfile.TemporaryDirectory() as tmp:
        ef_, tabs, g = eof_setup(tmp, eparams, kind, seed)
        sf, mf, ps = sl_setup(tmp, sparams, kind, seed + 1)
        xd, yd, zd, md = S.exponential_disc(ndisc, seed + 2, a=g['ascale'], h=g['hscale'])
        md = md * 0.025
        xh, yh, zh, mh = S.hernquist_halo(nhalo, seed + 3, a=ps['scale'], rmax=ps['rmax'])
        Pd = S.ParticleSet(xd, yd, zd, md)
        cosd, sind = eof.accumulate(Pd, tabs[0], tabs[4], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                                    g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'])
        H = particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = xh, yh, zh, mh
        coef = np.asarray(quiet(spheresl.compute_coefficients_solitary, H, sf, mf, verbose=0), dtype=np.float64)
        F = potential.Fields('none', ef_, sf, mf, verbose=0)
        F.EOF = eof.EOF_Object(); F.EOF.eof_file = ef_; F.EOF.cos = cosd; F.EOF.sin = sind
        F.SL = spheresl.SL_Object(); F.SL.sph_file = sf; F.SL.model_file = mf; F.SL.expcoef = coef
        F.halofac = 1.25
        quiet(F.prep_tables)
        F.set_field_parameters()
        # points: disc-like positions, a few generic ones
        px = np.concatenate([xd[:npts - 4], [0.02, -0.3, 1e-4, 0.004]])
        py = np.concatenate([yd[:npts - 4], [0.0, 0.4, -2e-4, 0.003]])
        pz = np.concatenate([zd[:npts - 4], [0.0, -0.2, 1e-5, 0.05]])
        rot = 0.83
        cart_full = np.array([[float(v) for v in F.return_forces_cart(px[i], py[i], pz[i], rotpos=rot)] for i in range(npts)])
        # Fields.return_forces_cyl (potential.py:389-440)
        cyl_full = np.array([[float(v) for v in F.return_forces_cyl(px[i], py[i], pz[i], rotpos=rot)] for i in range(npts)])
        F.set_field_parameters(no_odd=True, halo_l=2, halo_n=3, disk_m=2, disk_n=3)
        cart_trunc = np.array([[float(v) for v in F.return_forces_cart(px[i], py[i], pz[i], rotpos=-0.4)] for i in range(npts)])
        F.reset_field_parameters()
        # orbits
        vc = 1.0
        orbs = []
        pos0 = np.zeros((3, norb)); vel0 = np.zeros((3, norb))
        dt, rotfreq = 3.0e-4, -5.0
        for k in range(norb):
            R0 = 0.008 * (k + 1)
            a = F.return_forces_cart(R0, 0.0, 0.0)
            vcirc = np.sqrt(max(-R0 * float(a[0] + a[1]), 1e-12))
            pos0[:, k] = [R0, 0.0, 0.0005 * k]
            vel0[:, k] = [0.05 * vcirc, (0.7 + 0.1 * k) * vcirc, 0.02 * vcirc]
            O = integrate.leapfrog_integrate(F, nint, dt, pos0[:, k], vel0[:, k], rotfreq=rotfreq, force=True)
            orbs.append(np.array([O[key] for key in ('X', 'Y', 'Z', 'VX', 'VY', 'VZ', 'P', 'FX', 'FY', 'FZ', 'TX', 'TY', 'VTX', 'VTY', 'T')]))
        # one truncated / no_odd orbit, rotfreq > 0
        Ot = integrate.leapfrog_integrate(F, nint, dt, pos0[:, 0], vel0[:, 0], rotfreq=3.0, no_odd=True,
                                          halo_l=2, halo_n=4, disk_m=4, disk_n=5)
        orb_trunc = np.array([Ot[key] for key in ('X', 'Y', 'Z', 'VX', 'VY', 'VZ', 'P', 'TX', 'TY', 'VTX', 'VTY', 'T')])
        ts = np.array([integrate.compute_timestep(F, pos0[:, k], vel0[:, k]) for k in range(norb)])
        # orbit grids (integrate.py:760-922) and the orbit-map text format (513-537)
        F.reset_field_parameters()
        vc0 = np.sqrt(max(-0.01 * float(sum(F.return_forces_cart(0.01, 0.0, 0.0)[:2])), 1e-12))
        rads = np.array([0.006, 0.012, 0.02]); vels = np.array([0.6, 0.9]) * vc0
        grid = quiet(integrate.integrate_grid, rads, vels, F, 40, 1.0e-5, rotfreq, False, -1, -1, 100., 1000, 0)
        grid_y = quiet(integrate.integrate_grid_launchy, rads[:2], vels[:1], F, 30, 1.0e-5, 3.0, False, -1, -1, 50., 1000, 0)
        zs = np.array([0.0, 0.002]); vzs = np.array([0.05 * vc0])
        grid3 = quiet(integrate.integrate_grid_3D, rads[:2], vels, F, 24, 1.0e-5, rotfreq, False, -1, -1, 100., 1000, zs, vzs, 0)
        buf = io.StringIO(); integrate.print_orbit_array(buf, grid[:1, :1, :, :5]); orbit_txt = buf.getvalue()
        # frozen-field file (potential.py:738-1040): hash of the reference-written bytes + forces after restore
        import hashlib
        F.time = 0.25; F.xcen_disk = F.ycen_disk = F.zcen_disk = 0.; F.xcen_halo = F.ycen_halo = F.zcen_halo = 0.
        F.filename = 'snap'; F.eof_file = 'eof.cache'; F.sph_file = 'sl.cache'; F.model_file = 'sl.model'
        ff = os.path.join(tmp, 'frozen.field')
        F.save_field(ff)
        with open(ff, 'rb') as fh:
            field_sha = hashlib.sha256(fh.read()).hexdigest()
        field_size = os.path.getsize(ff)
        FR = potential.restore_field(ff)
        FR.set_field_parameters()
        cart_restored = np.array([[float(v) for v in FR.return_forces_cart(px[i], py[i], pz[i], rotpos=rot)] for i in range(npts)])
    np.savez_compressed(os.path.join(HERE, name + '.npz'),
                        meta=json.dumps(dict(eof_params=eparams, sl_params=sparams, kind=kind, seed=seed, geo=g,
                                             halofac=1.25, rot_full=rot, rot_trunc=-0.4, dt=dt, rotfreq=rotfreq, nint=nint)),
                        cos=cosd, sin=sind, coef=coef, px=px, py=py, pz=pz, cart_full=cart_full, cart_trunc=cart_trunc, cyl_full=cyl_full,
                        pos0=pos0, vel0=vel0, orbits=np.array(orbs), orbit_trunc=orb_trunc, timestep=ts,
                        grid_rads=rads, grid_vels=vels, grid=grid, grid_y=grid_y, grid_zs=zs, grid_vzs=vzs, grid3=grid3,
                        orbit_txt=np.array(orbit_txt), field_sha=np.array(field_sha), field_size=np.array(field_size),
                        cart_restored=cart_restored)
    print('wrote', name)


def case_eof_density(name, eparams, kind, seed, npart, rscale, zscale):
    """eof.accumulated_eval_particles(density=True, eof_file=...) (eof.py:1041-1050, 1106, 1122, 1136-1142) on a
    cache file written with dens=1: p0, p, d0, d, fr, fp, fz, R."""
    rng = np.random.default_rng(seed + 100)
    with tempfile.TemporaryDirectory() as tmp:
        f, tabs, g = eof_setup(tmp, eparams, kind, seed)
        potC, rfC, zfC, dC, potS, rfS, zfS, dS = tabs
        x, y, z, m = edge_particles(rng, npart, rscale, zscale)
        P = S.ParticleSet(x, y, z, m)
        cosd, sind = eof.accumulate(P, potC, potS, g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                                    g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'])
        full = quiet(eof.accumulated_eval_particles, P, cosd, sind, density=True, eof_file=f, verbose=0)
        win = quiet(eof.accumulated_eval_particles, P, cosd, sind, m1=1, m2=2, density=True, eof_file=f, verbose=0)
        assert len(full) == 8
    np.savez_compressed(os.path.join(HERE, name + '.npz'),
                        meta=json.dumps(dict(eof_params=eparams, kind=kind, seed=seed, geo=g)),
                        x=x, y=y, z=z, m=m, cos=cosd, sin=sind, full=np.array(full), win12=np.array(win))
    print('wrote', name)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'density':
    case_eof_density('eof_dens_random', dict(SMALL_EOF, cmap=1, dens=1), 'random', 14, 300, 0.02, 0.002)
    case_eof_density('eof_dens_smooth', dict(mmax=4, numx=48, numy=32, nmax=8, norder=6, dens=1), 'smooth', 15, 300, 0.015, 0.001)
elif __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'field':
    case_field('field_small', dict(mmax=4, numx=48, numy=32, nmax=8, norder=6), dict(lmax=4, nmax=6, numr=400),
               'smooth', 31, 3000, 800, 48, 3, 250)
    case_field('field_std', {}, dict(lmax=6), 'smooth', 32, 4000, 600, 24, 2, 120)
elif __name__ == '__main__':
    # EOF: small random (adversarial, cmap 1 and 0), standard-geometry smooth
    case_eof('eof_small_random_cmap1', dict(SMALL_EOF, cmap=1), 'random', 11, 400, 0.02, 0.002)
    case_eof('eof_small_random_cmap0', dict(SMALL_EOF, cmap=0, rmax=2.0), 'random', 12, 400, 0.005, 0.002)
    case_eof('eof_std_smooth', {}, 'smooth', 13, 1500, 0.015, 0.001)
    # SL: small random (cmap 1 and 0), standard lmax=4 and lmax=6
    case_sl('sl_small_random_cmap1', dict(SMALL_SL, cmap=1), 'random', 21, 150)
    case_sl('sl_small_random_cmap0', dict(SMALL_SL, cmap=0), 'random', 22, 150)
    case_sl('sl_std_l4', dict(lmax=4), 'smooth', 23, 300)
    case_sl('sl_std_l6', dict(lmax=6), 'smooth', 24, 200)
    # combined field + orbits: small geometry (cheap), smooth tables
    case_field('field_small', dict(mmax=4, numx=48, numy=32, nmax=8, norder=6), dict(lmax=4, nmax=6, numr=400),
               'smooth', 31, 3000, 800, 48, 3, 250)
    case_field('field_std', {}, dict(lmax=6), 'smooth', 32, 4000, 600, 24, 2, 120)
