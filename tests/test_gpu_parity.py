"""
GPU parity tests (run with -m gpu on the B200): the CUDA path, called through
the C ABI (exptool_b200.ops -> libbfe.so), against (a) the golden vectors the
unmodified reference produced and (b) the oracle on larger seeded inputs.

Tolerances (BASELINE.json north_star): <= 1e-10 relative (max-norm) for FP64
coefficients and forces; orbit states <= 1e-8.
"""
import numpy as np
import pytest

from helpers import load_golden, relerr, eof_tables, sl_tables, eof_geo_args, build_field, O, S

pytestmark = pytest.mark.gpu

TOL = 1e-10
ORBIT_TOL = 1e-8

EOF_CASES = ['eof_small_random_cmap1', 'eof_small_random_cmap0', 'eof_std_smooth']
SL_CASES = ['sl_small_random_cmap1', 'sl_small_random_cmap0', 'sl_std_l4', 'sl_std_l6']
FIELD_CASES = ['field_small', 'field_std']


@pytest.fixture(scope='module')
def ops():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU tests selected but CUDA is not available')
    from exptool_b200 import ops as _ops
    return _ops


def make_eof(ops, T, g):
    return ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                         g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                         rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])


def make_sl(ops, p, ev, ef, xi, p0, d0):
    return ops.SLTables(p['lmax'], p['nmax'], p['numr'], p['cmap'], p['scale'], ev, ef, xi, p0, d0)


@pytest.fixture(params=['direct', 'sorted', 'sorted_lane'])
def eof_mode(ops, request):
    """run the EOF tests through both kernel families (bfe_set_option); 'sorted' evaluates the sorted set on
    the FP64 tensor cores (default), 'sorted_lane' with the per-lane kernel"""
    v = 1 if request.param == 'direct' else 2
    ops.set_option('eof_accumulate_mode', v)
    ops.set_option('eof_force_mode', v)
    ops.set_option('force_mma', 0 if request.param == 'sorted_lane' else 1)
    yield 'direct' if v == 1 else 'sorted'
    ops.set_option('eof_accumulate_mode', 0)
    ops.set_option('eof_force_mode', 0)
    ops.set_option('force_mma', 1)


@pytest.mark.parametrize('name', EOF_CASES)
def test_eof_accumulate_golden(ops, eof_mode, name):
    d, meta = load_golden(name)
    p, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    c, s = E.accumulate(d['x'], d['y'], d['z'], d['m'])
    assert relerr(c.cpu().numpy(), d['cos']) < TOL
    assert relerr(s.cpu().numpy(), d['sin']) < TOL
    # repeat: the handle's workspace/counter must be reusable, and BOTH formulations are bit-deterministic (the sorted
    # one since round 2: stable, atomic-free tile counting sort -- ascending particle index inside every cell)
    c2, s2 = E.accumulate(d['x'], d['y'], d['z'], d['m'])
    assert np.array_equal(c2.cpu().numpy(), c.cpu().numpy()) and np.array_equal(s2.cpu().numpy(), s.cpu().numpy())


@pytest.mark.parametrize('name', EOF_CASES)
def test_eof_force_golden(ops, eof_mode, name):
    d, meta = load_golden(name)
    p, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    nf = meta['nforce']
    E.contract(d['cos'], d['sin'])
    out = E.force(d['x'][:nf], d['y'][:nf], d['z'][:nf]).cpu().numpy()
    for i in range(6):
        assert relerr(out[i], d['full'][i]) < TOL, i
    E.contract(d['cos'], d['sin'], m1=1, m2=2)
    out = E.force(d['x'][:nf], d['y'][:nf], d['z'][:nf]).cpu().numpy()
    for i in range(6):
        assert relerr(out[i], d['win12'][i]) < TOL, i
    # eof.force_eval variants at cylindrical points
    variants = dict(full=(g['mmax'], g['norder'], False), trunc=(max(g['mmax'] - 1, 1), max(g['norder'] - 1, 1), False),
                    noodd=(g['mmax'], g['norder'], True))
    for vname, (M, N, no_odd) in variants.items():
        E.contract(d['cos'], d['sin'], m1=0, m2=M, nuse=N, no_odd=no_odd)
        out = E.force_eval_points(d['pt_r'], d['pt_z'], d['pt_phi']).cpu().numpy()
        ref = d['fe_' + vname]
        for i in range(5):
            assert relerr(out[i], ref[:, i]) < TOL, (vname, i)


@pytest.fixture(params=['direct', 'sorted'])
def sl_mode(ops, request):
    ops.set_option('sl_accumulate_mode', 1 if request.param == 'direct' else 2)
    yield request.param
    ops.set_option('sl_accumulate_mode', 0)


@pytest.mark.parametrize('name', SL_CASES)
def test_sl_accumulate_golden(ops, sl_mode, name):
    d, meta = load_golden(name)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = make_sl(ops, p, ev, ef, xi, p0, d0)
    c = H.accumulate(d['x'], d['y'], d['z'], d['m']).cpu().numpy()
    assert relerr(c, d['coef']) < TOL
    c2 = H.accumulate(d['x'], d['y'], d['z'], d['m'], no_odd=True).cpu().numpy()
    assert relerr(c2, d['coef_noodd']) < TOL


@pytest.mark.parametrize('name', SL_CASES)
def test_sl_force_golden(ops, name):
    d, meta = load_golden(name)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = make_sl(ops, p, ev, ef, xi, p0, d0)
    nf = meta['nforce']
    xs, ys, zs = d['x'][1:nf + 1], d['y'][1:nf + 1], d['z'][1:nf + 1]
    for key, kw in (('allp', {}), ('allp_win12', dict(l1=1, l2=2)), ('allp_noodd', dict(no_odd=True))):
        H.contract(d['coef'], **kw)
        out = H.force(xs, ys, zs).cpu().numpy()
        ref = d[key]          # den0,den1,pot0,pot1,potr,pott,potp,rr
        for i, j in enumerate((2, 3, 4, 5, 6, 7)):
            assert relerr(out[i], ref[j]) < TOL, (key, j)
        H.contract_density(d['coef'], **kw)
        den = H.density(xs, ys, zs).cpu().numpy()
        for j in range(2):
            assert relerr(den[j], ref[j]) < TOL, (key, 'den', j)
    L, N = p['lmax'], p['nmax']
    for key, (l, n, no_odd) in dict(fe_full=(L, N, False), fe_trunc=(max(L - 1, 1), max(N - 2, 1), False),
                                    fe_noodd=(L, N, True)).items():
        H.contract(d['coef'], l1=0, l2=l, nuse=n, no_odd=no_odd)
        out = H.force_eval_points(d['pt_r'], d['pt_costh'], d['pt_phi'], trig_index_l=True).cpu().numpy()
        for i in range(5):
            assert relerr(out[i], d[key][:, i]) < TOL, (key, i)
    H.contract(d['coef'])
    out = H.force_eval_points(d['pt_r'], d['pt_costh'], d['pt_phi'], trig_index_l=False).cpu().numpy()
    for i, j in enumerate((4, 5, 6, 3, 2)):     # ae_full: den0,den1,pot0,pot1,potr,pott,potp
        assert relerr(out[i], d['ae_full'][:, j]) < TOL, j
    H.contract_density(d['coef'])
    den = H.density_eval_points(d['pt_r'], d['pt_costh'], d['pt_phi']).cpu().numpy()
    for j in range(2):
        assert relerr(den[j], d['ae_full'][:, j]) < TOL, ('den', j)


def _field_handles(ops, meta, d):
    pe, T, g = eof_tables(meta)
    ps, ev, ef, xi, p0, d0 = sl_tables(meta, seed_offset=1)
    return make_eof(ops, T, g), make_sl(ops, ps, ev, ef, xi, p0, d0), g, ps


@pytest.mark.parametrize('name', FIELD_CASES)
def test_field_cart_golden(ops, name):
    d, meta = load_golden(name)
    E, H, g, ps = _field_handles(ops, meta, d)
    E.contract(d['cos'], d['sin'])
    H.contract(meta['halofac'] * d['coef'])
    out = ops.field_force_cart(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_full']).cpu().numpy()
    for i in range(8):
        assert relerr(out[i], d['cart_full'][:, i]) < TOL, i
    E.contract(d['cos'], d['sin'], m1=0, m2=2, nuse=3, no_odd=True)
    H.contract(meta['halofac'] * d['coef'], l1=0, l2=2, nuse=3, no_odd=True)
    out = ops.field_force_cart(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_trunc']).cpu().numpy()
    for i in range(8):
        assert relerr(out[i], d['cart_trunc'][:, i]) < TOL, i


@pytest.mark.parametrize('name', FIELD_CASES)
def test_leapfrog_golden(ops, name):
    d, meta = load_golden(name)
    E, H, g, ps = _field_handles(ops, meta, d)
    E.contract(d['cos'], d['sin'])
    H.contract(meta['halofac'] * d['coef'])
    nint = meta['nint']
    state, traj, nsteps = ops.leapfrog(E, H, d['pos0'], d['vel0'], nint, meta['dt'], rotfreq=meta['rotfreq'],
                                       traj_stride=1)
    traj = traj.cpu().numpy()
    ref = d['orbits']        # (norb, 15, nint): X Y Z VX VY VZ P FX FY FZ ...
    assert (nsteps.cpu().numpy() == nint).all()
    for k in range(ref.shape[0]):
        for j in range(10):
            assert relerr(traj[:, j, k], ref[k, j]) < ORBIT_TOL, (k, j)
    # end state equals last trajectory sample
    assert np.array_equal(state.cpu().numpy(), traj[-1, :6, :])
    # truncated, no_odd, rotfreq > 0
    E.contract(d['cos'], d['sin'], m1=0, m2=4, nuse=5, no_odd=True)
    H.contract(meta['halofac'] * d['coef'], l1=0, l2=2, nuse=4, no_odd=True)
    state, traj, nsteps = ops.leapfrog(E, H, d['pos0'][:, :1], d['vel0'][:, :1], nint, meta['dt'], rotfreq=3.0,
                                       traj_stride=1)
    traj = traj.cpu().numpy()
    for j in range(7):
        assert relerr(traj[:, j, 0], d['orbit_trunc'][j]) < ORBIT_TOL, j


FP32_TABLE_TOL = 1e-5      # BASELINE.json north_star: "<= 1e-5 where FP32 table interpolation is used"


@pytest.mark.parametrize('name', FIELD_CASES)
def test_fp32_table_mode(ops, name):
    """Option table_fp32: the contracted tables of the per-point field kernels are stored as float (half the
    L1 tag cycles); everything else stays FP64.  Same goldens, the tolerance north_star states for this mode."""
    d, meta = load_golden(name)
    E, H, g, ps = _field_handles(ops, meta, d)
    E.contract(d['cos'], d['sin'])
    H.contract(meta['halofac'] * d['coef'])
    ops.set_option('table_fp32', 1)
    try:
        out = ops.field_force_cart(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_full']).cpu().numpy()
        errs = [relerr(out[i], d['cart_full'][:, i]) for i in range(8)]
        assert max(errs) < FP32_TABLE_TOL, errs
        assert max(errs) > 1e-12, 'FP32 tables not in use?'
        nint = meta['nint']
        state, traj, nsteps = ops.leapfrog(E, H, d['pos0'], d['vel0'], nint, meta['dt'], rotfreq=meta['rotfreq'],
                                           traj_stride=1)
        traj = traj.cpu().numpy()
        ref = d['orbits']
        for k in range(ref.shape[0]):
            for j in range(6):          # phase-space coordinates after nint steps of accumulated table rounding
                assert relerr(traj[:, j, k], ref[k, j]) < 10 * FP32_TABLE_TOL, (k, j)
        if ps['lmax'] in (4, 6):
            sl = H.force(d['px'], d['py'], d['pz']).cpu().numpy()
            ops.set_option('table_fp32', 0)
            sl64 = H.force(d['px'], d['py'], d['pz']).cpu().numpy()
            for i in range(5):
                assert relerr(sl[i], sl64[i]) < FP32_TABLE_TOL, i
    finally:
        ops.set_option('table_fp32', 0)
    # back in FP64 mode the contraction is re-expanded and parity is the FP64 one again
    out = ops.field_force_cart(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_full']).cpu().numpy()
    for i in range(8):
        assert relerr(out[i], d['cart_full'][:, i]) < TOL, i


# ---------------------------------------------------------------------------
# larger seeded inputs against the oracle (sizes the oracle finishes in seconds)
# ---------------------------------------------------------------------------
def test_eof_accumulate_force_oracle_50k(ops, eof_mode):
    p, T = S.make_eof_tables({}, kind='smooth')
    XMIN, XMAX, dX, YMIN, YMAX, dY = O.eof_set_table_params(RMAX=p['rmax'], RMIN=p['rmin'], ASCALE=p['ascale'],
                                                            HSCALE=p['hscale'], NUMX=p['numx'], NUMY=p['numy'], CMAP=p['cmap'])
    g = dict(XMIN=XMIN, dX=dX, YMIN=YMIN, dY=dY, numx=p['numx'], numy=p['numy'], mmax=p['mmax'], norder=p['norder'],
             ascale=p['ascale'], hscale=p['hscale'], cmap=p['cmap'])
    E = make_eof(ops, T, g)
    x, y, z, m = S.exponential_disc(50001, 2002)
    c, s = E.accumulate(x, y, z, m)
    co, so = O.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g['mmax'], g['norder'], *eof_geo_args(g),
                              g['ascale'], g['hscale'], g['cmap'])
    assert relerr(c.cpu().numpy(), co) < TOL
    assert relerr(s.cpu().numpy(), so) < TOL
    E.contract(c, s)
    out = E.force(x[:20000], y[:20000], z[:20000]).cpu().numpy()
    ref = O.eof_force_particles(x[:20000], y[:20000], z[:20000], co, so, T['potC'], T['rforceC'], T['zforceC'],
                                T['potS'], T['rforceS'], T['zforceS'], *eof_geo_args(g), g['mmax'], g['norder'],
                                g['ascale'], g['hscale'], g['cmap'])
    for i in range(6):
        assert relerr(out[i], ref[i]) < TOL, i
    # linearity (size-independent property): coefficients of a doubled-mass set are doubled,
    # coefficients of a concatenated set are the sum
    c2, s2 = E.accumulate(x, y, z, 2.0 * m)
    assert relerr(c2.cpu().numpy(), 2.0 * co) < TOL
    ca, sa = E.accumulate(x[:30000], y[:30000], z[:30000], m[:30000])
    cb, sb = E.accumulate(x[30000:], y[30000:], z[30000:], m[30000:])
    assert relerr((ca + cb).cpu().numpy(), co) < TOL
    # empty input
    c0, s0 = E.accumulate(x[:0], y[:0], z[:0], m[:0])
    assert float(c0.abs().max()) == 0.0 and float(s0.abs().max()) == 0.0


def test_eof_prepared_set_matches_separate_calls(ops):
    """bfe_eof_prepare + *_prepared (one cell sort for both passes) == the separate entry points."""
    meta = dict(eof_params={}, kind='smooth', seed=0)
    p, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    x, y, z, m = S.exponential_disc(40000, 31)
    x[:3] = [0.0, 5.0, -1e-9]; y[:3] = [0.0, 0.0, 1e-9]; z[:3] = [0.0, 1.0, -3.0]     # origin, far outside the table
    ops.set_option('eof_accumulate_mode', 1); ops.set_option('eof_force_mode', 1)
    c1, s1 = E.accumulate(x, y, z, m)
    E.contract(c1, s1)
    f1 = E.force(x, y, z).cpu().numpy()
    ops.set_option('eof_accumulate_mode', 0); ops.set_option('eof_force_mode', 0)
    E.prepare(x, y, z, m)
    c2, s2 = E.accumulate_prepared()
    assert relerr(c2.cpu().numpy(), c1.cpu().numpy()) < 1e-12
    assert relerr(s2.cpu().numpy(), s1.cpu().numpy()) < 1e-12
    E.contract(c1, s1)
    f2 = E.force_prepared().cpu().numpy()
    for i in range(6):
        assert relerr(f2[i], f1[i]) < 1e-12, i
    # a set prepared without masses cannot be accumulated
    E.prepare(x, y, z)
    with pytest.raises(Exception):
        E.accumulate_prepared()
    f3 = E.force_prepared().cpu().numpy()
    assert np.array_equal(f3[5], f2[5])


def test_eof_sorted_cell_edges(ops):
    """Particles placed ON and within 1e-4..1e-15 (relative) of the table's cell edges, on the z = 0 plane, at the
    origin and outside the table: the histogram pass indexes cells in FP32 with an FP64 fallback near edges, the
    scatter pass in FP64 -- any disagreement between the two would misfile records.  Sorted == direct == oracle."""
    meta = dict(eof_params={}, kind='smooth', seed=0)
    p, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    rng = np.random.default_rng(77)
    kx = rng.integers(0, g['numx'] + 1, 60000).astype(np.float64)
    ky = rng.integers(0, g['numy'] + 1, 60000).astype(np.float64)
    eps = 10.0 ** rng.uniform(-15.5, -4, 60000) * rng.choice([-1.0, 0.0, 1.0], 60000)
    on_x = rng.random(60000) < 0.6            # which coordinate sits on an edge
    X = np.where(on_x, kx, kx + rng.random(60000))
    Y = np.where(on_x, ky + rng.random(60000), ky)
    xi = g['XMIN'] + X * g['dX']
    yy = g['YMIN'] + Y * g['dY']
    if g['cmap'] == 1:
        xi = np.clip(xi, -0.999999, 0.999999)
        r = g['ascale'] * (1.0 + xi) / (1.0 - xi)
    elif g['cmap'] == 2:
        r = np.exp(xi)
    else:
        r = xi
    zz = g['hscale'] * np.sinh(yy)
    r = np.abs(r * (1.0 + np.where(on_x, eps, 0.0)))
    zz = zz * (1.0 + np.where(on_x, 0.0, eps))
    phi = rng.uniform(0, 2 * np.pi, 60000)
    x = r * np.cos(phi); y = r * np.sin(phi); z = zz.copy()
    z[:2000] = rng.choice([0.0, 1e-300, -1e-300, 1e-12, -1e-9, 1e-8], 2000)       # the y = 0 edge between two rows
    x[2000:2010] = 0.0; y[2000:2010] = 0.0                                         # axis
    x[2010:2020] *= 1e3; z[2020:2030] *= 1e3                                       # far outside the table
    m = rng.uniform(0.5, 1.5, 60000) / 60000
    ops.set_option('eof_accumulate_mode', 1); ops.set_option('eof_force_mode', 1)
    c1, s1 = E.accumulate(x, y, z, m)
    E.contract(c1, s1)
    f1 = E.force(x, y, z).cpu().numpy()
    ops.set_option('eof_accumulate_mode', 2); ops.set_option('eof_force_mode', 2)
    try:
        c2, s2 = E.accumulate(x, y, z, m)
        E.contract(c1, s1)
        f2 = E.force(x, y, z).cpu().numpy()
    finally:
        ops.set_option('eof_accumulate_mode', 0); ops.set_option('eof_force_mode', 0)
    assert relerr(c2.cpu().numpy(), c1.cpu().numpy()) < 1e-12
    assert relerr(s2.cpu().numpy(), s1.cpu().numpy()) < 1e-12
    for i in range(6):
        assert relerr(f2[i], f1[i]) < 1e-12, i
    co, so = O.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g['mmax'], g['norder'], *eof_geo_args(g),
                              g['ascale'], g['hscale'], g['cmap'])
    assert relerr(c1.cpu().numpy(), co) < TOL and relerr(s1.cpu().numpy(), so) < TOL


@pytest.mark.parametrize('lmax', [4, 6])
def test_sl_accumulate_force_oracle_30k(ops, sl_mode, lmax):
    meta = dict(sl_params=dict(lmax=lmax), kind='smooth', seed=0)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = make_sl(ops, p, ev, ef, xi, p0, d0)
    x, y, z, m = S.hernquist_halo(30011, 1001)
    c = H.accumulate(x, y, z, m)
    co = O.sl_accumulate(x, y, z, m, p['lmax'], p['nmax'], ev, ef, xi, p0, p['cmap'], p['scale'])
    assert relerr(c.cpu().numpy(), co) < TOL
    H.contract(c)
    out = H.force(x[:10000], y[:10000], z[:10000]).cpu().numpy()
    ref = O.sl_all_eval_particles(x[:10000], y[:10000], z[:10000], co, p['lmax'], p['nmax'], ev, ef, xi, p0, d0,
                                  p['cmap'], p['scale'])
    for i in range(6):
        assert relerr(out[i], ref[i]) < TOL, i
    c0 = H.accumulate(x[:0], y[:0], z[:0], m[:0])
    assert float(c0.abs().max()) == 0.0


def test_generic_kernels_large_basis(ops):
    """mmax > 6 / lmax > 6 take the generic (BFE_MAX_*) kernel instances."""
    meta = dict(eof_params=dict(mmax=8, numx=20, numy=12, nmax=8, norder=5), kind='random', seed=5)
    p, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    rng = np.random.default_rng(9)
    x = rng.normal(0, 0.02, 3000); y = rng.normal(0, 0.02, 3000); z = rng.normal(0, 0.002, 3000); m = rng.random(3000)
    c, s = E.accumulate(x, y, z, m)
    co, so = O.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g['mmax'], g['norder'], *eof_geo_args(g),
                              g['ascale'], g['hscale'], g['cmap'])
    assert relerr(c.cpu().numpy(), co) < TOL and relerr(s.cpu().numpy(), so) < TOL
    E.contract(c, s)
    out = E.force(x, y, z).cpu().numpy()
    ref = O.eof_force_particles(x, y, z, co, so, T['potC'], T['rforceC'], T['zforceC'], T['potS'], T['rforceS'],
                                T['zforceS'], *eof_geo_args(g), g['mmax'], g['norder'], g['ascale'], g['hscale'], g['cmap'])
    for i in range(6):
        assert relerr(out[i], ref[i]) < TOL, i
    meta = dict(sl_params=dict(lmax=8, nmax=6, numr=300), kind='random', seed=6)
    ps, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = make_sl(ops, ps, ev, ef, xi, p0, d0)
    xh, yh, zh, mh = S.hernquist_halo(2000, 77)
    ch = H.accumulate(xh, yh, zh, mh)
    cho = O.sl_accumulate(xh, yh, zh, mh, ps['lmax'], ps['nmax'], ev, ef, xi, p0, ps['cmap'], ps['scale'])
    assert relerr(ch.cpu().numpy(), cho) < TOL
    H.contract(ch)
    out = H.force(xh, yh, zh).cpu().numpy()
    ref = O.sl_all_eval_particles(xh, yh, zh, cho, ps['lmax'], ps['nmax'], ev, ef, xi, p0, d0, ps['cmap'], ps['scale'])
    for i in range(6):
        assert relerr(out[i], ref[i]) < TOL, i


# ---------------------------------------------------------------------------
# BASELINE.json's full sizes (configs[0], configs[1]) against the C restatement of the oracle
# ---------------------------------------------------------------------------
def test_config_c2_eof_1e6_accumulate_force(ops):
    """configs[1]: EOF mmax=6 norder=18, accumulate + force eval on a 10^6-particle exponential disc."""
    from oracle import oracle_c as OC
    meta = dict(eof_params={}, kind='smooth', seed=0)
    p, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    x, y, z, m = S.exponential_disc(1000000, 2002)
    co, so = OC.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g)
    fo = OC.eof_force(x, y, z, co, so, T, g)
    for mode in (1, 2):                       # direct and cell-sorted kernel families
        ops.set_option('eof_accumulate_mode', mode); ops.set_option('eof_force_mode', mode)
        c, s = E.accumulate(x, y, z, m)
        assert relerr(c.cpu().numpy(), co) < TOL and relerr(s.cpu().numpy(), so) < TOL, mode
        E.contract(co, so)
        f = E.force(x, y, z).cpu().numpy()
        for i in range(6):
            assert relerr(f[i], fo[i]) < TOL, (mode, i)
    ops.set_option('eof_accumulate_mode', 0); ops.set_option('eof_force_mode', 0)
    # the benchmark's own call sequence: one cell sort for both passes
    E.prepare(x, y, z, m)
    c, s = E.accumulate_prepared()
    E.contract(c, s)
    f = E.force_prepared().cpu().numpy()
    assert relerr(c.cpu().numpy(), co) < TOL
    for i in range(6):
        assert relerr(f[i], fo[i]) < TOL, i
    # size-independent properties: linearity in the masses, additivity over a split of the set
    c2, s2 = E.accumulate(x, y, z, 3.0 * m)
    assert relerr(c2.cpu().numpy(), 3.0 * co) < TOL
    ca, sa = E.accumulate(x[:400000], y[:400000], z[:400000], m[:400000])
    cb, sb = E.accumulate(x[400000:], y[400000:], z[400000:], m[400000:])
    assert relerr((ca + cb).cpu().numpy(), co) < TOL and relerr((sa + sb).cpu().numpy(), so) < TOL
    # a permutation of the particles changes only the summation order
    perm = np.random.default_rng(1).permutation(x.size)
    cp, sp = E.accumulate(x[perm], y[perm], z[perm], m[perm])
    assert relerr(cp.cpu().numpy(), co) < 1e-12


def test_config_c1_sl_1e5_accumulate_force(ops):
    """configs[0]: SL lmax=4 nmax=18, accumulation + force eval on a 10^5-particle Hernquist halo."""
    from oracle import oracle_c as OC
    meta = dict(sl_params=dict(lmax=4), kind='smooth', seed=0)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = make_sl(ops, p, ev, ef, xi, p0, d0)
    x, y, z, m = S.hernquist_halo(100000, 1001)
    co = OC.sl_accumulate(x, y, z, m, p['lmax'], p['nmax'], ev, ef, xi, p0, p['cmap'], p['scale'])
    for mode in (1, 2):
        ops.set_option('sl_accumulate_mode', mode)
        c = H.accumulate(x, y, z, m)
        assert relerr(c.cpu().numpy(), co) < TOL, mode
    ops.set_option('sl_accumulate_mode', 0)
    H.contract(co)
    f = H.force(x[:20000], y[:20000], z[:20000]).cpu().numpy()
    ref = O.sl_all_eval_particles(x[:20000], y[:20000], z[:20000], co, p['lmax'], p['nmax'], ev, ef, xi, p0, d0,
                                  p['cmap'], p['scale'])
    for i in range(6):
        assert relerr(f[i], ref[i]) < TOL, i
    c2 = H.accumulate(x, y, z, 2.0 * m)
    assert relerr(c2.cpu().numpy(), 2.0 * co) < TOL


def test_peer_allreduce_protocol_single_gpu(ops):
    """bfe_peer_allreduce (the coefficient sum over NVLink peer memory) with three "ranks" emulated on ONE GPU: three
    exchange buffers in this process, one bfe_peer per rank over the same three pointers, the three kernels launched on
    three streams (they must be co-resident: one CTA each).  200 back-to-back sums exercise the sequence / parity
    protocol; the result must be the rank-order sum, bit-identical for every rank, with no timeout."""
    import ctypes as C
    import torch
    from exptool_b200 import _lib
    lib = _lib.load()
    W, NMAX, N, ITERS = 3, 512, 252, 200
    bufs = (C.c_void_p * W)()
    for r in range(W):
        p = C.c_void_p(); handle = C.create_string_buffer(64)
        _lib.check(lib.bfe_peer_buffer_create(NMAX, C.byref(p), handle))
        bufs[r] = p.value
    peers = []
    for r in range(W):
        h = C.c_void_p()
        _lib.check(lib.bfe_peer_create(r, W, NMAX, bufs, C.byref(h)))
        peers.append(h)
    gen = torch.Generator(device='cuda'); gen.manual_seed(3)
    data = torch.randn(W, ITERS, N, dtype=torch.float64, device='cuda', generator=gen)
    want = torch.zeros(ITERS, N, dtype=torch.float64, device='cuda')
    for r in range(W):                       # rank order, as the kernel sums
        want += data[r]
    streams = [torch.cuda.Stream() for _ in range(W)]
    torch.cuda.synchronize()
    for it in range(ITERS):
        for r in range(W):
            _lib.check(lib.bfe_peer_allreduce(peers[r], C.c_void_p(data[r, it].data_ptr()), N,
                                              C.c_void_p(streams[r].cuda_stream)))
    torch.cuda.synchronize()
    for r in range(W):
        v = C.c_uint64(1)
        _lib.check(lib.bfe_peer_error(peers[r], C.c_void_p(0), C.byref(v)))
        assert v.value == 0, 'rank %d timed out at sequence %d' % (r, v.value)
        assert torch.equal(data[r], want), r
    # argument checks
    assert lib.bfe_peer_allreduce(peers[0], C.c_void_p(data[0, 0].data_ptr()), NMAX + 1, C.c_void_p(0)) != 0
    # a failed collective never hands out numbers: rank 1's buffer is marked failed (fault injection) -> its next call
    # returns NaN at once and raises the error word of every rank, so the ranks waiting for it give up with NaN too
    # (round 1 summed whatever the slots held after a 20 s time-out)
    _lib.check(lib.bfe_peer_poison(peers[1], 777, C.c_void_p(0)))
    blocks = torch.ones(W, N, dtype=torch.float64, device='cuda')
    t0 = __import__('time').time()
    for r in (1, 0, 2):
        _lib.check(lib.bfe_peer_allreduce(peers[r], C.c_void_p(blocks[r].data_ptr()), N, C.c_void_p(streams[r].cuda_stream)))
    torch.cuda.synchronize()
    assert __import__('time').time() - t0 < 10.0, 'the peers of a failed rank must not sit out the 20 s time-out'
    assert bool(torch.isnan(blocks).all())
    for r in range(W):
        v = C.c_uint64(0)
        _lib.check(lib.bfe_peer_error(peers[r], C.c_void_p(0), C.byref(v)))
        assert v.value != 0, r
    for h in peers:
        lib.bfe_peer_destroy(h)
    for r in range(W):
        _lib.check(lib.bfe_peer_buffer_destroy(C.c_void_p(bufs[r])))


def test_degenerate_sizes_every_entry_point(ops):
    """Empty, single and ragged (n = 33: one warp + 1) inputs through every evaluation entry point, in each per-point
    kernel variant: shapes are kept, nothing is written out of bounds, and n = 1 / n = 33 equal the same points taken
    from a larger call (the kernels have no dependence on the launch size)."""
    import torch
    meta = dict(eof_params=dict(mmax=4, numx=24, numy=16, nmax=8, norder=5), sl_params=dict(lmax=4, nmax=6, numr=200),
                kind='smooth', seed=3)
    pe, T, g = eof_tables(meta)
    ps, ev, ef, xi, p0, d0 = sl_tables(meta, seed_offset=1)
    E, H = make_eof(ops, T, g), make_sl(ops, ps, ev, ef, xi, p0, d0)
    x, y, z, m = S.exponential_disc(200, 5)
    c, s = E.accumulate(x, y, z, m)
    ch = H.accumulate(x, y, z, m)
    E.contract(c, s); H.contract(ch); H.contract_density(ch)
    r = np.sqrt(x * x + y * y) + 1e-6
    cth = np.clip(z / np.sqrt(r * r + z * z), -1, 1)
    phi = np.arctan2(y, x)
    try:
        for staged, blk, f32 in ((0, 0, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1)):
            ops.set_option('staged_eval', staged); ops.set_option('blk_eval', blk); ops.set_option('table_fp32', f32)
            ops.set_option('eof_force_mode', 1)
            full = dict(ef=E.force(x, y, z), ep=E.force_eval_points(r, z, phi), sf=H.force(x, y, z),
                        sp=H.force_eval_points(r, cth, phi, trig_index_l=True),
                        fc=ops.field_force_cart(E, H, x, y, z, rotpos=0.4), fy=ops.field_force_cyl(E, H, x, y, z, rotpos=0.4),
                        dn=H.density(x, y, z), dp=H.density_eval_points(r, cth, phi))
            for n in (0, 1, 33):
                sl = slice(0, n)
                got = dict(ef=E.force(x[sl], y[sl], z[sl]), ep=E.force_eval_points(r[sl], z[sl], phi[sl]),
                           sf=H.force(x[sl], y[sl], z[sl]), sp=H.force_eval_points(r[sl], cth[sl], phi[sl], trig_index_l=True),
                           fc=ops.field_force_cart(E, H, x[sl], y[sl], z[sl], rotpos=0.4),
                           fy=ops.field_force_cyl(E, H, x[sl], y[sl], z[sl], rotpos=0.4),
                           dn=H.density(x[sl], y[sl], z[sl]), dp=H.density_eval_points(r[sl], cth[sl], phi[sl]))
                for k in full:
                    assert got[k].shape == (full[k].shape[0], n), (k, n)
                    assert torch.equal(got[k], full[k][:, :n]), (k, n, staged, blk, f32)
            st, tr, ns = ops.leapfrog(E, H, np.stack([x, y, z])[:, :0], np.zeros((3, 0)), 5, 1e-3)
            assert st.shape == (6, 0)
            st1, _, _ = ops.leapfrog(E, H, np.stack([x, y, z])[:, :1], np.zeros((3, 1)), 5, 1e-3)
            st33, _, _ = ops.leapfrog(E, H, np.stack([x, y, z])[:, :33], np.zeros((3, 33)), 5, 1e-3)
            assert torch.equal(st1[:, 0], st33[:, 0])
    finally:
        ops.set_option('staged_eval', 1); ops.set_option('blk_eval', 1); ops.set_option('table_fp32', 0)
        ops.set_option('eof_force_mode', 0)
    # building blocks on empty input
    Vc, Vs = E.get_pot(r[:0], z[:0])
    assert Vc.shape == (g['mmax'] + 1, g['norder'], 0)
    assert all(o.shape[-1] == 0 for o in H.radial_matrices(r[:0]))
    P, dP = ops.legendre_tables(4, cth[:0])
    assert P.shape == (5, 5, 0)


@pytest.mark.parametrize('lmax', [4, 6])
def test_sl_register_deposit_option(ops, lmax):
    """option sl_deposit_mode = 2 (lane-register run sums) gives the slab kernel's coefficients, incl. no_odd"""
    p, ev, ef, xi, p0, d0 = sl_tables(dict(sl_params=dict(lmax=lmax), kind='smooth', seed=0))
    H = make_sl(ops, p, ev, ef, xi, p0, d0)
    x, y, z, m = S.hernquist_halo(70001, 9)
    ops.set_option('sl_accumulate_mode', 2)
    try:
        for no_odd in (False, True):
            ops.set_option('sl_deposit_mode', 1)
            a = H.accumulate(x, y, z, m, no_odd=no_odd).cpu().numpy()
            ops.set_option('sl_deposit_mode', 2)
            b = H.accumulate(x, y, z, m, no_odd=no_odd).cpu().numpy()
            assert relerr(b, a) < 1e-13, (lmax, no_odd)
    finally:
        ops.set_option('sl_accumulate_mode', 0); ops.set_option('sl_deposit_mode', 0)


@pytest.mark.parametrize('name', FIELD_CASES)
def test_cell_sorted_field_and_leapfrog_paths(ops, name):
    """The cell-ordered paths for large point sets / orbit batches (bfe_orbit_sort.cu), forced onto the small golden
    cases: Fields.return_forces_cart/_cyl in chunks of 7 points, leapfrog re-sorted every 4 steps.  Same goldens, and
    bit-identical to the paths in the caller's order (the per-point arithmetic is the same function)."""
    import torch
    d, meta = load_golden(name)
    E, H, g, ps = _field_handles(ops, meta, d)
    E.contract(d['cos'], d['sin'])
    H.contract(meta['halofac'] * d['coef'])
    if ps['lmax'] not in (4, 6):
        pytest.skip('block kernels cover lmax 4 and 6')
    plain_c = ops.field_force_cart(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
    plain_y = ops.field_force_cyl(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
    nint = meta['nint']
    plain_s, _, plain_n = ops.leapfrog(E, H, d['pos0'], d['vel0'], nint, meta['dt'], rotfreq=meta['rotfreq'])
    dts = np.full(d['pos0'].shape[1], meta['dt']) * (1.0 + 0.1 * np.arange(d['pos0'].shape[1]))
    plain_d, _, _ = ops.leapfrog(E, H, d['pos0'], d['vel0'], nint, dts, rotfreq=meta['rotfreq'])
    saved = {k: ops.get_option(k) for k in ('field_sort_min', 'field_sort_chunk', 'orbit_sort_min', 'orbit_resort', 'table_fp32')}
    try:
        ops.set_option('field_sort_min', 1); ops.set_option('field_sort_chunk', 7)
        ops.set_option('orbit_sort_min', 1); ops.set_option('orbit_resort', 4)
        for f32 in (0, 1):
            ops.set_option('table_fp32', f32)
            sc = ops.field_force_cart(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
            sy = ops.field_force_cyl(E, H, d['px'], d['py'], d['pz'], rotpos=meta['rot_full'])
            ss, _, sn = ops.leapfrog(E, H, d['pos0'], d['vel0'], nint, meta['dt'], rotfreq=meta['rotfreq'])
            sd, _, _ = ops.leapfrog(E, H, d['pos0'], d['vel0'], nint, dts, rotfreq=meta['rotfreq'])
            if f32 == 0:
                assert torch.equal(sc, plain_c) and torch.equal(sy, plain_y)
                assert torch.equal(ss, plain_s) and torch.equal(sn, plain_n) and torch.equal(sd, plain_d)
                out = sc.cpu().numpy()
                for i in range(8):
                    assert relerr(out[i], d['cart_full'][:, i]) < TOL, i
                st = ss.cpu().numpy()
                for k in range(d['orbits'].shape[0]):
                    for j in range(6):
                        assert abs(st[j, k] - d['orbits'][k, j, -1]) <= ORBIT_TOL * np.max(np.abs(d['orbits'][k, j])), (k, j)
            else:
                assert relerr(sc.cpu().numpy(), plain_c.cpu().numpy()) < FP32_TABLE_TOL
                assert relerr(ss.cpu().numpy(), plain_s.cpu().numpy()) < 10 * FP32_TABLE_TOL
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)


@pytest.mark.parametrize('lmax', [4, 6])
def test_key_ordered_paths_at_size(ops, lmax):
    """The key-ordered evaluation (composite (EOF cell, SL interval) key, two chunks in flight on two
    streams; bfe_orbit_sort.cu) on 1.5e5 mixed disc + halo points in uneven chunks, and the re-sorted leapfrog on 3e4 orbits:
    bit-identical to the caller-order kernels (same per-point arithmetic), for re-sort intervals 1, 3 and per-orbit steps."""
    import torch
    meta = dict(eof_params={}, sl_params=dict(lmax=lmax), kind='smooth', seed=0)
    pe, T, g = eof_tables(meta)
    ps, ev, ef, xi, p0, d0 = sl_tables(meta, seed_offset=1)
    E, H = make_eof(ops, T, g), make_sl(ops, ps, ev, ef, xi, p0, d0)
    xd, yd, zd, md = S.exponential_disc(90000, 11)
    xh, yh, zh, mh = S.hernquist_halo(60001, 12)
    c, s_ = E.accumulate(xd, yd, zd, md)
    E.contract(c * 0.025, s_ * 0.025)
    H.contract(H.accumulate(xh, yh, zh, mh))
    x = np.concatenate([xd, xh]); y = np.concatenate([yd, yh]); z = np.concatenate([zd, zh])
    keys = ('field_sort_min', 'field_sort_chunk', 'orbit_sort_min', 'orbit_resort', 'stage_eval')
    saved = {k: ops.get_option(k) for k in keys}
    try:
        ops.set_option('field_sort_min', 0); ops.set_option('orbit_resort', 0)
        ref_c = ops.field_force_cart(E, H, x, y, z, rotpos=0.3)
        ref_y = ops.field_force_cyl(E, H, x, y, z, rotpos=-1.1)
        norb, nint = 30001, 26
        pos0 = np.stack([xd[:norb], yd[:norb], zd[:norb]])
        R = np.sqrt(pos0[0] ** 2 + pos0[1] ** 2) + 1e-9
        vel0 = np.stack([-pos0[1] / R, pos0[0] / R, 0.05 * np.cos(np.arange(norb))]) * 1.3
        dts = 3e-4 * (1.0 + 0.5 * np.sin(np.arange(norb)))
        ref_s, _, ref_n = ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=-5.0)
        ref_d, _, _ = ops.leapfrog(E, H, pos0, vel0, nint, dts, rotfreq=2.0)
        ops.set_option('field_sort_min', 1); ops.set_option('field_sort_chunk', 40000); ops.set_option('orbit_sort_min', 1)
        def same(a, b, stage):
            # stage 0 (per-lane loads, the default): the same arithmetic as the caller-order kernels -> the same bits.
            # stage 1 (table blocks of every tile staged in shared memory by TMA bulk copies, option stage_eval) reads the
            # three-node SL blocks A3, the per-lane kernels the per-interval polynomial blocks A4: equal to rounding.
            if stage == 0:
                return torch.equal(a, b)
            a2, b2 = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)          # per output row, the parity tolerance
            return bool(((a2 - b2).abs().amax(dim=1) <= TOL * b2.abs().amax(dim=1)).all())
        for stage in (0, 1):
            ops.set_option('stage_eval', stage)
            assert same(ops.field_force_cart(E, H, x, y, z, rotpos=0.3), ref_c, stage), stage
            assert same(ops.field_force_cyl(E, H, x, y, z, rotpos=-1.1), ref_y, stage), stage
            for K in (1, 3):
                ops.set_option('orbit_resort', K)
                st, _, ns = ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=-5.0)
                assert same(st, ref_s, stage) and torch.equal(ns, ref_n), (stage, K)
                st, _, _ = ops.leapfrog(E, H, pos0, vel0, nint, dts, rotfreq=2.0)
                assert same(st, ref_d, stage), (stage, K)
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)


def test_stable_cell_sort_is_bit_reproducible(ops):
    """The stable, atomic-free tile counting sort behind bfe_eof_prepare (option sort_stable, default): the sorted accumulate
    and the sorted force evaluation are bit-identical from run to run (round 1: slot claims by global integer atomics, order
    inside a cell = arrival order, coefficients reproducible to ~1e-16 only), equal to the atomic-claim path and to the direct
    kernel within rounding, for ragged sizes, a set concentrated in ONE cell, and a set with NaNs."""
    import torch
    meta = dict(eof_params={}, kind='smooth', seed=0)
    pe, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    saved = {k: ops.get_option(k) for k in ('sort_stable', 'eof_accumulate_mode', 'eof_force_mode')}
    try:
        for n, seed in ((300001, 5), (40000, 6), (1025, 7)):
            x, y, z, m = S.exponential_disc(n, seed)
            if seed == 6:                                   # all particles in one table cell: the longest possible group
                x = 0.01 + 1e-7 * x; y = 1e-7 * y; z = 1e-6 + 1e-9 * z
            if seed == 7:
                x = x.copy(); x[3] = np.nan
            ops.set_option('eof_accumulate_mode', 2); ops.set_option('eof_force_mode', 2)
            ops.set_option('sort_stable', 1)
            runs = []
            for rep in range(3):
                c, s_ = E.accumulate(x, y, z, m)
                E.contract(c, s_)
                runs.append((c.clone(), s_.clone(), E.force(x, y, z).clone()))
            for rep in (1, 2):
                for a, b in zip(runs[0], runs[rep]):
                    assert torch.equal(a, b) or (seed == 7 and torch.equal(torch.isnan(a), torch.isnan(b))), (n, rep)
            if seed == 7:
                continue
            ops.set_option('sort_stable', 0)
            c0, s0 = E.accumulate(x, y, z, m)
            assert relerr(c0.cpu().numpy(), runs[0][0].cpu().numpy()) < 1e-12 and relerr(s0.cpu().numpy(), runs[0][1].cpu().numpy()) < 1e-12
            ops.set_option('eof_accumulate_mode', 1)
            c1, s1 = E.accumulate(x, y, z, m)
            assert relerr(c1.cpu().numpy(), runs[0][0].cpu().numpy()) < 1e-12
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)


@pytest.mark.gpu
@pytest.mark.parametrize('lmax', [4, 6])
def test_stable_sl_bin_sort_is_bit_reproducible(ops, lmax):
    """The sorted SL accumulation with the stable tile sort of the radial bins and the static task assignment of the deposit
    kernel (option sort_stable, default): the coefficients are bit-identical from run to run, for ragged sizes and a set
    concentrated in one radial interval; equal within rounding to the path with slot claims by integer atomics + the dynamic
    task queue, and to the direct kernel."""
    import torch
    meta = dict(sl_params=dict(lmax=lmax), kind='smooth', seed=0)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = make_sl(ops, p, ev, ef, xi, p0, d0)
    saved = {k: ops.get_option(k) for k in ('sort_stable', 'sl_accumulate_mode')}
    try:
        for n, seed in ((300001, 11), (50000, 12), (1025, 13)):
            x, y, z, m = S.hernquist_halo(n, seed)
            if seed == 12:                                  # one radial interval: the longest possible group
                rr = np.sqrt(x * x + y * y + z * z)
                f = (0.02 * (1.0 + 1e-7 * np.arange(n) / n)) / rr
                x, y, z = x * f, y * f, z * f
            ops.set_option('sl_accumulate_mode', 2)
            ops.set_option('sort_stable', 1)
            runs = [H.accumulate(x, y, z, m).clone() for _ in range(3)]
            assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2]), (n, lmax)
            ops.set_option('sort_stable', 0)
            c0 = H.accumulate(x, y, z, m)
            assert relerr(c0.cpu().numpy(), runs[0].cpu().numpy()) < 1e-12, (n, lmax)
            ops.set_option('sl_accumulate_mode', 1)
            c1 = H.accumulate(x, y, z, m)
            assert relerr(c1.cpu().numpy(), runs[0].cpu().numpy()) < 1e-12, (n, lmax)
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)


@pytest.mark.gpu
def test_key_subbits_change_relays_workspace(ops):
    """Option key_subbits changes the size of the key-sort workspace's header (nkeys = ncell << subbits).  Changing it between
    calls on the same handles must re-lay the workspace (round 2 bug: the old allocation was kept and the larger histogram
    overran the item arrays -> CUDA error / silent corruption).  Results are the caller-order bits for every setting."""
    import torch
    meta = dict(eof_params={}, sl_params=dict(lmax=6), kind='smooth', seed=0)
    pe, T, g = eof_tables(meta)
    E = make_eof(ops, T, g)
    p, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = make_sl(ops, p, ev, ef, xi, p0, d0)
    xd, yd, zd, md = S.exponential_disc(120000, 21)
    xh, yh, zh, mh = S.hernquist_halo(80000, 22)
    c, s_ = E.accumulate(xd, yd, zd, md)
    E.contract(c * 0.025, s_ * 0.025)
    H.contract(H.accumulate(xh, yh, zh, mh))
    x = np.concatenate([xd, xh]); y = np.concatenate([yd, yh]); z = np.concatenate([zd, zh])
    keys = ('field_sort_min', 'field_sort_chunk', 'key_subbits', 'orbit_key_subbits', 'key_mode', 'orbit_sort_min', 'orbit_resort')
    saved = {k: ops.get_option(k) for k in keys}
    try:
        ops.set_option('key_mode', 0)                 # the bit scheme: its key count follows the two options
        ops.set_option('field_sort_min', 0); ops.set_option('orbit_resort', 0)
        ref = ops.field_force_cart(E, H, x, y, z, rotpos=0.3)
        pos0 = np.stack([xd[:20000], yd[:20000], zd[:20000]])
        vel0 = np.stack([-pos0[1], pos0[0], 0.0 * pos0[2]]) * 1.1
        ref_s, _, _ = ops.leapfrog(E, H, pos0, vel0, 14, 3e-4, rotfreq=-5.0)
        ops.set_option('field_sort_min', 1); ops.set_option('field_sort_chunk', 70000)
        ops.set_option('orbit_sort_min', 1); ops.set_option('orbit_resort', 3)
        for sub, osub in ((4, 4), (8, 4), (2, 8), (0, 0), (6, 7), (7, 4)):
            ops.set_option('key_subbits', sub); ops.set_option('orbit_key_subbits', osub)
            assert torch.equal(ops.field_force_cart(E, H, x, y, z, rotpos=0.3), ref), (sub, osub)
            st, _, _ = ops.leapfrog(E, H, pos0, vel0, 14, 3e-4, rotfreq=-5.0)
            assert torch.equal(st, ref_s), (sub, osub)
            ops.set_option('key_mode', 3 - ops.get_option('key_mode'))          # alternate with the per-cell spans (points and orbits): another key count again
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)
