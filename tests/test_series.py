"""Coefficient time-series consumers: eof.reorganize_eof_dict, eof.calculate_eof_phase and the two
signal helpers, against outputs of the unmodified reference (tests/golden/eof_series.npz).  Host post-processing: no GPU."""
import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))
from helpers import GOLDEN      # noqa: E402

from exptool_b200.basis import eof          # noqa: E402
from exptool_b200.utils import utils        # noqa: E402


def _series():
    d = np.load(os.path.join(GOLDEN, 'eof_series.npz'))
    D = {}
    for i in range(d['t'].size):
        o = eof.EOF_Object()
        o.time = d['t'][i]; o.mmax = d['cos'].shape[1] - 1; o.nmax = d['cos'].shape[2]
        o.cos = d['cos'][i].copy(); o.sin = d['sin'][i].copy()
        D[i] = o
    return d, D


def _flatten(d, prefix=''):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(_flatten(v, prefix + str(k) + '__'))
        else:
            out[prefix + str(k)] = np.asarray(v, dtype=np.float64)
    return out


def _check(got, d, tag, tol=1e-12):
    flat = _flatten(got)
    keys = [k for k in d.files if k.startswith(tag)]
    assert len(keys) == len(flat) and len(keys) > 0
    for k in keys:
        a, b = flat[k[len(tag):]], d[k]
        assert a.shape == b.shape, k
        assert np.array_equal(np.isnan(a), np.isnan(b)), k
        if np.isfinite(b).any():
            assert np.nanmax(np.abs(a - b)) <= tol * max(np.nanmax(np.abs(b)), 1.0), k


def test_reorganize_eof_dict_matches_reference():
    d, D = _series()
    _check(eof.reorganize_eof_dict(D), d, 'reorg__')


def test_calculate_eof_phase_matches_reference():
    d, D = _series()
    with contextlib.redirect_stdout(io.StringIO()):
        _check(eof.calculate_eof_phase(D, filter=False), d, 'phase__')
        d, D = _series()
        _check(eof.calculate_eof_phase(D, filter=False, nonan=True), d, 'phase_nonan__')


def test_pattern_speed_of_a_rotating_series():
    """size-independent property: a pattern rotating at a constant rate gives that rate as the speed (filtered)."""
    nt, mmax, nmax, omega = 400, 2, 3, 7.5
    D = {}
    for i in range(nt):
        o = eof.EOF_Object()
        o.time = 0.01 * i; o.mmax = mmax; o.nmax = nmax
        ph = omega * o.time * np.arange(mmax + 1)[:, None] * np.ones((1, nmax))
        o.cos = np.cos(ph); o.sin = np.sin(ph); o.sin[0] = 0
        D[i] = o
    with contextlib.redirect_stdout(io.StringIO()):
        DC = eof.calculate_eof_phase(D, filter=True, smooth_box=21, smooth_order=2)
    for mm in (1, 2):
        assert np.allclose(DC['netspeed'][mm][30:-30], mm * omega, rtol=1e-6)
        assert np.allclose(DC['speed'][mm][30:-30, 0], mm * omega, rtol=1e-6)
        assert DC['direction'][mm][0] > 0.9


def test_savitzky_golay_reproduces_polynomials_and_scipy():
    x = np.linspace(-1, 1, 101)
    y = 3 * x ** 2 - x + 0.5
    assert np.allclose(utils.savitzky_golay(y, 11, 2)[6:-6], y[6:-6], atol=1e-12)
    from scipy.signal import savgol_filter
    rng = np.random.default_rng(0)
    z = np.cumsum(rng.standard_normal(300))
    assert np.allclose(utils.savitzky_golay(z, 15, 3)[8:-8], savgol_filter(z, 15, 3)[8:-8], atol=1e-10)


def test_unwrap_phase_equals_reference_live():
    """where the reference is mounted: utils.unwrap_phase on random series, both winding directions, in-place +pi shift"""
    if not os.path.isdir('/root/reference/exptool'):
        import pytest
        pytest.skip('reference not mounted')
    import importlib
    from oracle import refshim
    refshim.load()
    rutils = importlib.import_module('exptool.utils.utils')
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(2, 120))
        t = np.sort(rng.random(n))
        if trial % 3 == 0:
            ph = rng.uniform(-np.pi, np.pi, n)
        elif trial % 3 == 1:
            w = rng.uniform(1, 40) * (1 if trial % 2 else -1)
            ph = np.arctan2(np.sin(7.3 * t * w), np.cos(7.3 * t * w))
        else:
            ph = rng.uniform(0, 2 * np.pi, n)
        a, b = ph.copy(), ph.copy()
        with contextlib.redirect_stdout(io.StringIO()):
            ra, rb = rutils.unwrap_phase(t, a), utils.unwrap_phase(t, b)
        assert np.array_equal(a, b) and np.array_equal(ra, rb), trial
