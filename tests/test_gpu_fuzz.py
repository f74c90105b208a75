"""Randomised consistency sweep over sizes, particle distributions and runtime options (GPU): the key-ordered / sorted
formulations against the caller-order / direct kernels they must reproduce.  Bounded (~20 s); seeds are fixed."""
import os

import numpy as np
import pytest

from helpers import eof_tables, sl_tables
from exptool_b200 import synthetic as S

pytestmark = pytest.mark.gpu

# BFE_FUZZ_ITERS / BFE_FUZZ_SEED scale the sweep up for a one-off hunt (defaults keep the suite at a few seconds)
ITERS = int(os.environ.get('BFE_FUZZ_ITERS', '0'))
SEED = int(os.environ.get('BFE_FUZZ_SEED', '0'))


@pytest.fixture(scope='module')
def ops():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU tests selected but CUDA is not available')
    from exptool_b200 import ops as _ops
    return _ops


def _handles(ops, lmax):
    meta = dict(eof_params={}, sl_params=dict(lmax=lmax), kind='smooth', seed=0)
    p, T, g = eof_tables(meta)
    E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'],
                      g['ascale'], g['hscale'], g['cmap'], rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'],
                      zforceS=T['zforceS'])
    ps, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
    return E, H


def _points(rng, n):
    kind = rng.integers(0, 6)
    if kind == 0:
        x, y, z, m = S.exponential_disc(n, int(rng.integers(1, 10**6)))
    elif kind == 1:
        x, y, z, m = S.hernquist_halo(n, int(rng.integers(1, 10**6)))
    elif kind == 2:                                       # everything in one table cell / one radial interval
        x = 0.013 + 1e-9 * rng.standard_normal(n); y = 1e-9 * rng.standard_normal(n); z = 2e-4 + 1e-10 * rng.standard_normal(n)
        m = np.full(n, 1.0 / n)
    elif kind == 3:                                       # on the axis, in the plane, far outside the tables
        x = rng.standard_normal(n) * 0.02; y = rng.standard_normal(n) * 0.02; z = rng.standard_normal(n) * 0.01
        k = max(n // 5, 1)
        x[:k] = 0.0; y[:k] = 0.0
        z[k:2 * k] = 0.0
        x[2 * k:3 * k] *= 1e4
        m = rng.uniform(0.5, 1.5, n) / n
    elif kind == 4:                                       # a thin ring: few cells, many intervals shared
        phi = rng.uniform(0, 2 * np.pi, n); R = 0.03 * (1 + 1e-3 * rng.standard_normal(n))
        x, y, z = R * np.cos(phi), R * np.sin(phi), 1e-3 * rng.standard_normal(n)
        m = np.full(n, 1.0 / n)
    else:                                                 # log-uniform radii over six decades
        r = 10.0 ** rng.uniform(-5, 1, n); mu = rng.uniform(-1, 1, n); phi = rng.uniform(0, 2 * np.pi, n)
        s = np.sqrt(1 - mu * mu)
        x, y, z = r * s * np.cos(phi), r * s * np.sin(phi), r * mu
        m = rng.uniform(0.5, 1.5, n) / n
    return [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, m)]


@pytest.mark.parametrize('lmax', [4, 6])
def test_fuzz_key_ordered_paths(ops, lmax):
    import torch
    E, H = _handles(ops, lmax)
    rng = np.random.default_rng(1234 + lmax + SEED)
    xd, yd, zd, md = S.exponential_disc(100000, 3)
    xh, yh, zh, mh = S.hernquist_halo(100000, 4)
    c, s_ = E.accumulate(xd, yd, zd, md)
    E.contract(c * 0.025, s_ * 0.025)
    H.contract(H.accumulate(xh, yh, zh, mh))
    keys = ('field_sort_min', 'field_sort_chunk', 'key_subbits', 'orbit_key_subbits', 'key_mode', 'field_eval_static', 'field_support_slim', 'field_gather_stream', 'orbit_sort_min', 'orbit_resort', 'pdl')
    saved = {k: ops.get_option(k) for k in keys}
    try:
        for it in range(ITERS or 20):
            n = int(rng.choice([1, 31, 129, 4097, 70001, int(rng.integers(2, 250000))]))
            x, y, z, _ = _points(rng, n)
            rot = float(rng.uniform(-3, 3))
            ops.set_option('field_sort_min', 0); ops.set_option('orbit_resort', 0)
            ref_c = ops.field_force_cart(E, H, x, y, z, rotpos=rot)
            ref_y = ops.field_force_cyl(E, H, x, y, z, rotpos=rot)
            norb = min(n, 20000)
            pos0 = np.stack([x[:norb], y[:norb], z[:norb]])
            vel0 = np.stack([-pos0[1], pos0[0], 0.1 * pos0[2]]) * float(rng.uniform(0.5, 3.0))
            nint = int(rng.integers(9, 30))
            dts = 3e-4 * (1.0 + 0.5 * rng.random(norb))
            rf = float(rng.uniform(-6, 6)) if it % 2 else 0.0
            ref_s, _, ref_n = ops.leapfrog(E, H, pos0, vel0, nint, dts, rotfreq=rf)
            ops.set_option('field_sort_min', 1); ops.set_option('orbit_sort_min', 1)
            ops.set_option('field_sort_chunk', int(rng.choice([1000, 65536, 1 << 20, 1 << 22])))
            ops.set_option('key_subbits', int(rng.integers(0, 9))); ops.set_option('orbit_key_subbits', int(rng.integers(0, 9)))
            ops.set_option('orbit_resort', int(rng.integers(1, 6))); ops.set_option('pdl', int(rng.integers(0, 2)))
            ops.set_option('key_mode', int(rng.integers(0, 4))); ops.set_option('field_eval_static', int(rng.integers(0, 2))); ops.set_option('field_support_slim', int(rng.integers(0, 3)))
            ops.set_option('field_gather_stream', int(rng.integers(0, 2)))
            assert torch.equal(ops.field_force_cart(E, H, x, y, z, rotpos=rot), ref_c), (it, n)
            assert torch.equal(ops.field_force_cyl(E, H, x, y, z, rotpos=rot), ref_y), (it, n)
            st, _, ns = ops.leapfrog(E, H, pos0, vel0, nint, dts, rotfreq=rf)
            assert torch.equal(st, ref_s) and torch.equal(ns, ref_n), (it, n, nint)
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)


def test_fuzz_sorted_accumulate_and_force(ops):
    import torch
    E, H = _handles(ops, 6)
    rng = np.random.default_rng(4321 + SEED)
    keys = ('sort_stable', 'eof_accumulate_mode', 'eof_force_mode', 'sl_accumulate_mode', 'pdl')
    saved = {k: ops.get_option(k) for k in keys}

    def rel(a, b):
        d = float((a - b).abs().max()); s = float(b.abs().max())
        return d / s if s > 0 else d
    try:
        for it in range(ITERS or 14):
            n = int(rng.choice([1, 33, 1025, 8193, 150001, int(rng.integers(2, 300000))]))
            x, y, z, m = _points(rng, n)
            ops.set_option('pdl', int(rng.integers(0, 2)))
            ops.set_option('eof_accumulate_mode', 1); ops.set_option('eof_force_mode', 1); ops.set_option('sl_accumulate_mode', 1)
            c1, s1 = E.accumulate(x, y, z, m)
            h1 = H.accumulate(x, y, z, m)
            E.contract(c1, s1)
            f1 = E.force(x, y, z)
            ops.set_option('eof_accumulate_mode', 2); ops.set_option('eof_force_mode', 2); ops.set_option('sl_accumulate_mode', 2)
            for stable in (1, 0):
                ops.set_option('sort_stable', stable)
                c2, s2 = E.accumulate(x, y, z, m)
                h2 = H.accumulate(x, y, z, m)
                E.contract(c1, s1)
                f2 = E.force(x, y, z)
                assert rel(c2, c1) < 1e-11 and rel(s2, s1) < 1e-11 and rel(h2, h1) < 1e-11, (it, n, stable)
                fscale = float(f1[2:5].abs().max())                 # the three force components share one scale (fp may cancel to ~0)
                for i in range(6):
                    if 2 <= i <= 4:
                        assert float((f2[i] - f1[i]).abs().max()) <= 1e-11 * fscale, (it, n, stable, i)
                    else:
                        assert rel(f2[i], f1[i]) < 1e-11, (it, n, stable, i)
                if stable:
                    c3, s3 = E.accumulate(x, y, z, m)
                    h3 = H.accumulate(x, y, z, m)
                    assert torch.equal(c3, c2) and torch.equal(s3, s2) and torch.equal(h3, h2), (it, n)
                    assert torch.equal(E.force(x, y, z), f2), (it, n)
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)


def test_fuzz_host_pipelines(ops):
    """The chunked host-array pipelines (EOF and SL, accumulate and evaluate) against the one-shot device path for random sizes,
    pinned / pageable / mixed inputs, copy-thread counts and the one-upload-per-snapshot option."""
    import torch
    E, H = _handles(ops, 6)
    rng = np.random.default_rng(777 + SEED)
    keys = ('host_reuse', 'host_threads')
    saved = {k: ops.get_option(k) for k in keys}

    def rel(a, b):
        a = np.asarray(a); b = np.asarray(b)
        s = np.abs(b).max()
        return np.abs(a - b).max() / s if s > 0 else np.abs(a - b).max()
    try:
        for it in range(ITERS or 8):
            n = int(rng.choice([1, 5, 4099, 250001, int(rng.integers(2, 900000))]))
            x, y, z, m = _points(rng, n)
            ops.set_option('host_reuse', int(rng.integers(0, 2))); ops.set_option('host_threads', int(rng.integers(1, 6)))
            mode = int(rng.integers(0, 3))                 # 0 pageable NumPy, 1 pinned tensors, 2 mixed
            def host(a, k):
                if mode == 1 or (mode == 2 and k % 2 == 0):
                    return torch.from_numpy(a).pin_memory()
                return a
            P = [host(a, k) for k, a in enumerate((x, y, z, m))]
            cd, sd = E.accumulate(x, y, z, m)
            hd = H.accumulate(x, y, z, m)
            ch, sh = E.accumulate_host(*P)
            hh = H.accumulate_host(*P)
            assert rel(ch, cd.cpu().numpy()) < 1e-11 and rel(sh, sd.cpu().numpy()) < 1e-11, (it, n, mode)
            assert rel(hh, hd.cpu().numpy()) < 1e-11, (it, n, mode)
            E.contract(cd, sd); H.contract(hd)
            fe = E.force(x, y, z).cpu().numpy(); fs = H.force(x, y, z).cpu().numpy()
            ge = E.force_host(*P[:3]); gs = H.force_host(*P[:3])
            es = np.abs(fe[2:5]).max(); ss = np.abs(fs[2:5]).max()
            for i in range(6):
                if 2 <= i <= 4:
                    assert np.abs(ge[i] - fe[i]).max() <= 1e-11 * es and np.abs(gs[i] - fs[i]).max() <= 1e-11 * ss, (it, n, mode, i)
                else:
                    assert rel(ge[i], fe[i]) < 1e-11 and rel(gs[i], fs[i]) < 1e-11, (it, n, mode, i)
    finally:
        for k, v in saved.items():
            ops.set_option(k, v)
