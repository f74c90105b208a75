"""
make_series_golden.py -- golden vectors for the coefficient time-series consumers:
eof.reorganize_eof_dict (eof.py:1920-1958) and eof.calculate_eof_phase (eof.py:1961-2108, filter=False: the
reference's filter=True path calls np.mat, which NumPy 2 removed), produced by the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_series_golden.py
"""
import contextlib
import copy
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim                      # noqa: E402

reof = refshim.load()['eof']


def synthetic_series(seed=5, mmax=3, nmax=4, nt=60):
    """a rotating pattern per (m, n) plus noise, one weak (below-threshold) channel, keys in shuffled time order"""
    rng = np.random.default_rng(seed)
    times = np.arange(nt) * 0.02
    perm = rng.permutation(nt)
    cos = np.zeros((nt, mmax + 1, nmax)); sin = np.zeros((nt, mmax + 1, nmax)); t = np.zeros(nt)
    for idx, k in enumerate(perm):
        ph = 2.0 * np.pi * (times[k] * (3.0 + np.arange(nmax)[None, :] * 0.7) * (np.arange(mmax + 1)[:, None] - 1.5))
        amp = 1.0 / (1 + np.arange(nmax))[None, :] * np.ones((mmax + 1, 1))
        amp[2, 3] = 1e-4
        cos[idx] = amp * np.cos(ph) + 0.01 * rng.standard_normal((mmax + 1, nmax))
        sin[idx] = amp * np.sin(ph) + 0.01 * rng.standard_normal((mmax + 1, nmax))
        sin[idx, 0] = 0
        t[idx] = times[k]
    return t, cos, sin


def as_dict(t, cos, sin, cls):
    D = {}
    for i in range(t.size):
        o = cls()
        o.time = t[i]; o.mmax = cos.shape[1] - 1; o.nmax = cos.shape[2]
        o.cos = cos[i].copy(); o.sin = sin[i].copy()
        D[i] = o
    return D


def flatten(d, prefix=''):
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(flatten(v, prefix + str(k) + '__'))
        else:
            out[prefix + str(k)] = np.asarray(v, dtype=np.float64)
    return out


def main():
    t, cos, sin = synthetic_series()
    D = as_dict(t, cos, sin, reof.EOF_Object)
    with contextlib.redirect_stdout(io.StringIO()):
        A = reof.calculate_eof_phase(copy.deepcopy(D), filter=False)
        A2 = reof.calculate_eof_phase(copy.deepcopy(D), filter=False, nonan=True)
        C = reof.reorganize_eof_dict(copy.deepcopy(D))
    out = dict(t=t, cos=cos, sin=sin)
    for tag, d in (('phase__', A), ('phase_nonan__', A2), ('reorg__', C)):
        out.update({tag + k: v for k, v in flatten(d).items()})
    np.savez_compressed(os.path.join(HERE, 'eof_series.npz'), **out)
    print('wrote eof_series', len(out), 'arrays')


if __name__ == '__main__':
    main()
