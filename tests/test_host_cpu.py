"""
CPU-only tests (no GPU, no compute calls into libbfe.so): the C-ABI library loads and
exports every symbol include/bfe.h declares; host-side readers / file formats; the
sharding logic of parallel.py under gloo with world_size 2; the product path fails loudly
without CUDA.
"""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from helpers import ROOT, S, O, relerr


def test_library_exports_every_declared_symbol():
    from exptool_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    with open(os.path.join(ROOT, 'include', 'bfe.h')) as f:
        declared = sorted(set(re.findall(r'\b(bfe_[a-z0-9_]+)\s*\(', f.read())))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, name
    for name in _lib.SIGNATURES:
        assert name in declared, 'binding for undeclared symbol ' + name
    # calls that need no GPU
    assert lib.bfe_version() >= 100
    assert lib.bfe_error_string(0) == b'ok'
    assert b'argument' in lib.bfe_error_string(-1)
    assert lib.bfe_launch_count() >= 0


def test_struct_layout_matches_header():
    """sizeof of the ctypes mirrors == sizeof of the C structs (compiled with gcc here)."""
    from exptool_b200 import _lib
    import ctypes as C
    src = ('#include <stdio.h>\n#include "bfe.h"\nint main(){printf("%zu %zu\\n", sizeof(bfe_eof_params), '
           'sizeof(bfe_sl_params));return 0;}\n')
    with tempfile.TemporaryDirectory() as tmp:
        cfile = os.path.join(tmp, 't.c')
        with open(cfile, 'w') as f:
            f.write(src)
        exe = os.path.join(tmp, 't')
        subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), cfile, '-o', exe])
        a, b = [int(v) for v in subprocess.check_output([exe]).split()]
    assert C.sizeof(_lib.EofParams) == a
    assert C.sizeof(_lib.SlParams) == b


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from exptool_b200 import ops
    with pytest.raises(RuntimeError, match='no CPU path'):
        ops.dev(np.zeros(3))
    from exptool_b200.basis import eof
    p, T = S.make_eof_tables(dict(mmax=1, numx=4, numy=4, nmax=2, norder=2), kind='random')
    P = S.ParticleSet(np.zeros(2), np.zeros(2), np.zeros(2), np.ones(2))
    with pytest.raises(RuntimeError):
        eof.accumulate(P, T['potC'], T['potS'], 1, 2, -1.0, 0.1, -1.0, 0.1, 4, 4, 0.01, 0.001, 1)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under exptool_b200/ may reference it."""
    pkg = os.path.join(ROOT, 'exptool_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert not re.search(r'^\s*(from|import)\s+oracle', txt, re.M), fn
                assert '/root/reference' not in txt, fn


def test_eof_cache_reader_and_geometry():
    from exptool_b200.basis import eof
    for dens in (0, 1):
        p, T = S.make_eof_tables(dict(mmax=3, numx=10, numy=6, nmax=8, norder=4, dens=dens), kind='random', seed=3)
        with tempfile.TemporaryDirectory() as tmp:
            f = S.write_eof_cache(os.path.join(tmp, 'c'), p, T)
            rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, d = eof.eof_params(f)
            assert (numx, numy, mmax, norder, cmap, d) == (10, 6, 3, 4, 1, dens)
            tabs = eof.parse_eof(f)
            D = eof.read_eof_file(f)
        names = ('potC', 'rforceC', 'zforceC', 'densC', 'potS', 'rforceS', 'zforceS', 'densS')
        for name, a in zip(names, tabs):
            assert np.array_equal(a, T[name]), name
            assert np.array_equal(D[name], T[name])
    got = eof.set_table_params(RMAX=20.0, RMIN=0.001, ASCALE=0.01, HSCALE=0.001, NUMX=128, NUMY=64, CMAP=1)
    ref = O.eof_set_table_params(RMAX=20.0, RMIN=0.001, ASCALE=0.01, HSCALE=0.001, NUMX=128, NUMY=64, CMAP=1)
    assert np.allclose(np.array(got, dtype=float), np.array(ref), rtol=0, atol=0)
    # SURVEY.md App. A.1 numbers
    assert abs(float(got[0]) + 0.998002) < 1e-6 and abs(float(got[2]) - 0.0145775) < 1e-7
    assert abs(float(got[3]) + 5.644903) < 1e-6 and abs(float(got[5]) - 0.1764032) < 1e-7


def test_eof_new_style_header():
    """new-style cache: magic 0xc0a57a1, length, YAML (eof.py:126-157, 257-259)."""
    import yaml
    from exptool_b200.basis import eof
    p, T = S.make_eof_tables(dict(mmax=2, numx=6, numy=4, nmax=8, norder=3), kind='random', seed=4)
    hdr = yaml.safe_dump(dict(mmax=p['mmax'], numx=p['numx'], numy=p['numy'], nmax=p['nmax'], norder=p['norder'],
                              dens=False, cmap=1, rmin=p['rmin'], rmax=p['rmax'], ascl=p['ascale'], hscl=p['hscale'],
                              cmass=1.0, time=0.0)).encode()
    with tempfile.TemporaryDirectory() as tmp:
        old = S.write_eof_cache(os.path.join(tmp, 'old'), p, T)
        with open(old, 'rb') as f:
            body = f.read()[76:]
        new = os.path.join(tmp, 'new')
        with open(new, 'wb') as f:
            np.array([0xc0a57a1, len(hdr)], dtype='<i4').tofile(f)
            f.write(hdr)
            f.write(body)
        assert eof.eof_params(new) == eof.eof_params(old)
        for a, b in zip(eof.parse_eof(new), eof.parse_eof(old)):
            assert np.array_equal(a, b)


def test_sl_cache_reader_and_init_table():
    from exptool_b200.utils import halo_methods
    from helpers import hernquist_model_columns
    p, ev, ef = S.make_sl_tables(dict(lmax=3, nmax=5, numr=80), kind='random', seed=8)
    with tempfile.TemporaryDirectory() as tmp:
        sf = S.write_sl_cache(os.path.join(tmp, 's'), p, ev, ef)
        mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=p['scale'])
        lmax, nmax, numr, cmap, rmin, rmax, scale, ltable, evtable, eftable = halo_methods.read_cached_table(sf)
        assert halo_methods.parse_slgrid(sf)[:4] == (3, 5, 80, 1)
        xi, r, p0, d0 = halo_methods.init_table(mf, numr, rmin, rmax, cmap, scale)
    assert (lmax, nmax, numr, cmap) == (3, 5, 80, 1)
    assert np.array_equal(evtable, ev) and np.array_equal(eftable, ef)
    assert np.array_equal(ltable, np.arange(4))
    R1, D1, P1 = hernquist_model_columns(p['scale'])
    xo, ro, po, do = O.sl_init_table(R1, D1, P1, numr, rmin, rmax, cmap, scale)
    assert np.array_equal(xi, xo) and relerr(p0, po) < 1e-15 and relerr(d0, do) < 1e-15


def test_coefficient_dump_roundtrip_and_layout():
    from exptool_b200.basis import eof, spheresl
    rng = np.random.default_rng(0)
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, 'eofcoef')
        objs = []
        for k in range(3):
            E = eof.EOF_Object()
            E.time = np.float32(0.1 * k); E.filename = 'snap%d' % k; E.comp = 'star'; E.nbodies = 1000 + k
            E.eof_file = '.eof.cache'; E.mmax = 6; E.nmax = 18
            E.cos = rng.standard_normal((7, 18)); E.sin = rng.standard_normal((7, 18))
            eof.save_eof_coefficients(f, E)
            objs.append(E)
        # 4 + k*(16(mmax+1)nmax + 224)
        assert os.path.getsize(f) == 4 + 3 * (16 * 7 * 18 + 224)
        last, D = eof.restore_eof_coefficients(f)
        assert len(D) == 3 and last.nbodies == 1002
        for E, (t, R) in zip(objs, D.items()):
            assert np.array_equal(R.cos, E.cos) and np.array_equal(R.sin, E.sin) and R.mmax == 6 and R.nmax == 18
        f2 = os.path.join(tmp, 'slcoef')
        Sx = spheresl.SL_Object()
        Sx.time = np.float32(0.5); Sx.filename = 'snap'; Sx.comp = 'dark'; Sx.nbodies = 7
        Sx.sph_file = 'a'; Sx.model_file = 'b'; Sx.lmax = 4; Sx.nmax = 18
        Sx.expcoef = rng.standard_normal((25, 18))
        spheresl.save_sl_coefficients(f2, Sx)
        spheresl.save_sl_coefficients(f2, Sx)
        assert os.path.getsize(f2) == 4 + 2 * (8 * 25 * 18 + 324)      # App. B.4
        last, D = spheresl.restore_sl_coefficients(f2)
        assert np.array_equal(last.expcoef, Sx.expcoef) and last.lmax == 4
        # keyed by np.round(time, 3) like the reference (spheresl.py:1461): a caller can index with the plain number
        assert list(D.keys()) == [np.round(np.float32(0.5), 3)] and D[0.5] is last
        Sx.time = np.float32(0.2504)
        spheresl.save_sl_coefficients(f2, Sx)
        last, D = spheresl.restore_sl_coefficients(f2)
        assert np.round(np.float32(0.2504), 3) in D and len(D) == 2
        # a truncated dump raises (the reference does not swallow it, unlike eof.restore_eof_coefficients)
        with open(f2, 'r+b') as fh:
            fh.truncate(os.path.getsize(f2) - 100)
        with pytest.raises(Exception):
            spheresl.restore_sl_coefficients(f2)


def test_factorial_return_matches_oracle():
    """spheresl.factorial_return (spheresl.py:823-863) against the oracle's restatement (and, in test_oracle_vs_reference,
    the oracle against the reference itself)"""
    from exptool_b200.basis import spheresl
    for lmax in (0, 1, 4, 6, 10):
        a, b = spheresl.factorial_return(lmax), O.factorial_return(lmax)
        assert a.shape == (lmax + 1, lmax + 1) and np.allclose(a, b, rtol=1e-15, atol=0)
        assert np.all(np.triu(a, 1) == 0.0)


def test_torch_library_ops_registered_and_cuda_only():
    """torch.ops.exptool_b200.* exist with their schemas and have no CPU kernel (no fallback)."""
    import torch
    from exptool_b200 import torch_ops
    for name in torch_ops.OPS:
        op = getattr(torch.ops.exptool_b200, name)
        assert 'exptool_b200::' + name in str(op.default._schema)
    z = torch.zeros(4, dtype=torch.float64)
    with pytest.raises(NotImplementedError):
        torch.ops.exptool_b200.eof_force(0, z, z, z)
    with pytest.raises(NotImplementedError):
        torch.ops.exptool_b200.field_force_cart(0, 0, z, z, z, 0.0)


def test_shard_bounds_match_reference_partition():
    from exptool_b200 import parallel
    for n, w in ((10, 3), (1000003, 8), (7, 8), (0, 2), (16, 1)):
        b = parallel.shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        ref = O.partition_like_reference(n, w)
        assert [x[0] for x in b] + [n] == ref
        avg = n // w
        assert all(hi - lo == avg for lo, hi in b[1:])         # rank 0 takes the remainder


def test_particle_containers():
    from exptool_b200.io import particle
    x, y, z, m = (np.arange(3.0) + k for k in range(4))
    H = particle.holder(); H.xpos, H.ypos, H.zpos, H.mass = x, y, z, m
    P = particle.Particles(x, y, z, m)
    for c in (H, P, (x, y, z, m), S.ParticleSet(x, y, z, m)):
        a = particle.particle_arrays(c)
        assert all(np.array_equal(u, v) for u, v in zip(a, (x, y, z, m)))
    assert particle.particle_arrays((x, y, z))[3] is None
    with pytest.raises(TypeError):
        particle.particle_arrays(3)


GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import numpy as np, torch, torch.distributed as dist
from exptool_b200 import parallel, synthetic as S
from oracle import oracle_np as O
from helpers import eof_tables, sl_tables, eof_geo_args

class FakeEOF:            # stands in for ops.EOFTables: same accumulate() contract, oracle arithmetic on the host
    def __init__(s, T, g): s.T, s.g, s.mmax, s.norder = T, g, g['mmax'], g['norder']
    def accumulate(s, x, y, z, m):
        g = s.g
        c, q = O.eof_accumulate(np.asarray(x), np.asarray(y), np.asarray(z), np.asarray(m), s.T['potC'], s.T['potS'],
                                g['mmax'], g['norder'], *eof_geo_args(g), g['ascale'], g['hscale'], g['cmap'])
        return torch.from_numpy(c), torch.from_numpy(q)

class FakeSL:
    def __init__(s, p, ev, ef, xi, p0): s.a = (p, ev, ef, xi, p0); s.nrow = (p['lmax']+1)**2; s.nmax = p['nmax']
    def accumulate(s, x, y, z, m, no_odd=False):
        p, ev, ef, xi, p0 = s.a
        return torch.from_numpy(O.sl_accumulate(np.asarray(x), np.asarray(y), np.asarray(z), np.asarray(m), p['lmax'],
                                                p['nmax'], ev, ef, xi, p0, p['cmap'], p['scale'], no_odd=no_odd))

dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{port}', rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
meta = dict(eof_params=dict(mmax=2, numx=10, numy=8, nmax=8, norder=3), sl_params=dict(lmax=2, nmax=3, numr=50),
            kind='random', seed=1)
p, T, g = eof_tables(meta); E = FakeEOF(T, g)
ps, ev, ef, xi, p0, d0 = sl_tables(meta); H = FakeSL(ps, ev, ef, xi, p0)
x, y, z, m = S.exponential_disc(1001, 5)
# (1) global arrays replicated on both ranks -> each takes its block, one allreduce
c, s = parallel.eof_accumulate_sharded(E, x, y, z, m)
c0, s0 = E.accumulate(x, y, z, m)
assert float((c - c0).abs().max()) <= 1e-12 * float(c0.abs().max())
assert float((s - s0).abs().max()) <= 1e-12 * float(s0.abs().max())
lo, hi = parallel.my_shard(1001)
assert (lo, hi) == ((0, 501) if rank == 0 else (501, 1001))
# (2) already-sharded inputs (weak scaling) + the SL path
xs, ys, zs, ms = (a[lo:hi] for a in (x, y, z, m))
h = parallel.sl_accumulate_sharded(H, xs, ys, zs, ms, already_sharded=True)
h0 = H.accumulate(x, y, z, m)
assert float((h - h0).abs().max()) <= 1e-12 * float(h0.abs().max())
# (3) time series: 3 snapshots, ONE allreduce of the [3, ncoef] block
snaps = [((xs * (1 + 0.1 * k), ys, zs, ms), (xs, ys * (1 + 0.1 * k), zs, ms)) for k in range(3)]
cs, ss, hs = parallel.accumulate_series(E, H, snaps)
for k in range(3):
    ck, sk = E.accumulate(x * (1 + 0.1 * k), y, z, m)
    hk = H.accumulate(x, y * (1 + 0.1 * k), z, m)
    assert float((cs[k] - ck).abs().max()) <= 1e-12 * float(ck.abs().max())
    assert float((hs[k] - hk).abs().max()) <= 1e-12 * float(hk.abs().max())
# (4) coefficient broadcast
t = torch.full((4,), float(rank + 1), dtype=torch.float64)
parallel.broadcast_(t, src=0)
assert float(t.sum()) == 4.0
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_sharded_accumulate_gloo_world2():
    import socket
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as tmp:
        script = os.path.join(tmp, 'w.py')
        with open(script, 'w') as f:
            f.write(GLOO_WORKER.format(root=ROOT, port=port))
        procs = [subprocess.Popen([sys.executable, script, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                  text=True) for r in range(2)]
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=240)
            except subprocess.TimeoutExpired:
                p.kill()
                out, _ = p.communicate()
            outs.append(out)
        for r, (p, out) in enumerate(zip(procs, outs)):
            assert p.returncode == 0, out[-3000:]
            assert 'rank %d ok' % r in out


def _fields_from_golden(tmp, d, meta):
    """exptool_b200 Fields holding the golden coefficient set, built without a GPU (device handles are lazy)."""
    from exptool_b200.basis import potential
    pe, T = S.make_eof_tables(meta['eof_params'], kind=meta['kind'], seed=meta['seed'])
    ef = S.write_eof_cache(os.path.join(tmp, 'eof.cache'), pe, T)
    ps, ev, efn = S.make_sl_tables(meta['sl_params'], kind=meta['kind'], seed=meta['seed'] + 1)
    sf = S.write_sl_cache(os.path.join(tmp, 'sl.cache'), ps, ev, efn)
    mf = S.write_hernquist_model(os.path.join(tmp, 'sl.model'), a=ps['scale'])
    return potential.make_fields(ef, sf, mf, d['cos'], d['sin'], d['coef'], halofac=meta['halofac'])


@pytest.mark.parametrize('name', ['field_small', 'field_std'])
def test_frozen_field_file_is_byte_compatible(name):
    """Fields.save_field writes exactly the bytes the reference writes (sha256 recorded in the golden);
    restore_field reads them back with the float32-truncated geometry (potential.py:828-849)."""
    import hashlib
    from helpers import load_golden
    from exptool_b200.basis import potential
    d, meta = load_golden(name)
    with tempfile.TemporaryDirectory() as tmp:
        F = _fields_from_golden(tmp, d, meta)
        F.time = 0.25
        F.filename = 'snap'; F.eof_file = 'eof.cache'; F.sph_file = 'sl.cache'; F.model_file = 'sl.model'
        ff = os.path.join(tmp, 'frozen.field')
        F.save_field(ff)
        assert os.path.getsize(ff) == int(d['field_size'])
        with open(ff, 'rb') as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == str(d['field_sha'])
        R = potential.restore_field(ff)
    assert R.eof_file == b'eof.cache' and R.SL.sph_file == b'sl.cache' and int(R.nhalo) == 1000000
    assert (int(R.mmax), int(R.norder), int(R.lmaxhalo), int(R.nmaxhalo)) == (F.mmax, F.norder, F.lmaxhalo, F.nmaxhalo)
    assert np.array_equal(R.potC, F.potC) and np.array_equal(R.zforceS, F.zforceS) and np.array_equal(R.eftablehalo, F.eftablehalo)
    assert np.array_equal(R.EOF.cos, d['cos']) and np.array_equal(R.SL.expcoef, d['coef'])
    assert R.XMIN == np.float32(F.XMIN) and R.dY == np.float32(F.dY) and R.halofac == np.float32(meta['halofac'])
    assert abs(float(R.XMIN) - float(F.XMIN)) > 0          # the f4 truncation is real (App. C #15)


def test_orbit_map_text_format_roundtrip():
    from helpers import load_golden
    from exptool_b200.utils import integrate
    import io
    d, meta = load_golden('field_small')
    buf = io.StringIO()
    integrate.print_orbit_array(buf, d['grid'][:1, :1, :, :5])
    assert buf.getvalue() == str(d['orbit_txt'])              # same characters as the reference's writer
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, 'omap.txt')
        with open(fn, 'w') as f:
            integrate.print_orbit_array(f, d['grid'])
        D = integrate.read_integrations(fn)
        g = d['grid']
        assert len(D['X']) == g.shape[0] * g.shape[1]
        assert np.allclose(D['X'][3], g[1, 1, 0]) and np.allclose(D['VY'][3], g[1, 1, 5]) and D['dT'][3] == g[1, 1, 6, 1]
        fn3 = os.path.join(tmp, 'omap3.txt')
        with open(fn3, 'w') as f:
            integrate.print_orbit_array_3D(f, d['grid3'])
        D3 = integrate.read_integrations_3D(fn3)
        g3 = d['grid3']
        assert np.allclose(D3['Z'][1], g3[0, 0, 1, 0, 2]) and np.allclose(D3['VZ'][1], g3[0, 0, 1, 0, 7])


def test_table_cache_modes_and_invalidate():
    """eof.set_table_cache_mode: 'sampled' (default) can miss an in-place edit between sampled values -- invalidate_tables()
    is the explicit remedy -- 'full' hashes every byte."""
    from exptool_b200.basis import eof
    a = np.arange(200000.0).reshape(200, 1000)
    try:
        f1 = eof._fingerprint(a)
        a[0, 0] += 1.0                       # the first / last 64 values are always part of the tag
        assert eof._fingerprint(a) != f1
        eof.set_table_cache_mode('full')
        f1 = eof._fingerprint(a)
        a[57, 311] += 1.0
        assert eof._fingerprint(a) != f1
        with pytest.raises(ValueError):
            eof.set_table_cache_mode('sometimes')
    finally:
        eof.set_table_cache_mode('sampled')
    eof.invalidate_tables(a)                 # no entry yet: a no-op that must not raise
    eof.invalidate_tables()


def test_leapfrog_integrate_accepts_foreign_field_objects():
    """integrate.leapfrog_integrate with the reference's duck-typed FieldInstance protocol (integrate.py:61-64): an object
    that is not an exptool_b200 Fields is integrated with its own return_forces_cart -- here a Kepler potential, whose
    energy the velocity-Verlet scheme conserves to O(dt^2)."""
    from exptool_b200.utils import integrate

    class Kepler(object):
        def set_field_parameters(self, **kw):
            self.kw = kw

        def return_forces_cart(self, x, y, z, rotpos=0.0):
            r = np.sqrt(x * x + y * y + z * z)
            return -x / r ** 3, 0.0, -y / r ** 3, 0.0, -z / r ** 3, 0.0, -1.0 / r, 0.0
    K = Kepler()
    O = integrate.leapfrog_integrate(K, 400, 0.01, [1.0, 0.0, 0.0], [0.0, 0.9, 0.1], rotfreq=0.2, force=True, no_odd=True)
    assert K.kw['no_odd'] is True and len(O['T']) == 400 and 'FX' in O and 'TX' in O
    e = 0.5 * (O['VX'] ** 2 + O['VY'] ** 2 + O['VZ'] ** 2) + O['P']
    assert np.max(np.abs(e - e[0])) < 1e-4
    Oa = integrate.leapfrog_integrate(K, 4000, 0.01, [1.0, 0.0, 0.0], [0.0, 0.9, 0.1], apse=True, ap_max=2)
    assert 2 < len(Oa['T']) < 4000           # stopped after the second apocentre
    with pytest.raises(TypeError):
        integrate.leapfrog_integrate_batch(K, 10, 0.01, np.zeros((3, 2)), np.zeros((3, 2)))
