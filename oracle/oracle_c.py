"""
oracle_c.py -- ctypes loader for the C restatement (oracle/bfe_oracle.c).  TEST INFRASTRUCTURE ONLY:
used by tests/ (full-size parity) and by bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os

import numpy as np
from scipy.special import gammaln

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class EofGeom(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('mmax', 'norder', 'numx', 'numy', 'cmap')] + \
               [(k, C.c_double) for k in ('xmin', 'dx', 'ymin', 'dy', 'ascale', 'hscale')]


class SlGeom(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('lmax', 'nmax', 'numr', 'cmap')] + [('scale', C.c_double)]


def load():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, '_build', 'libbfe_oracle.so')
        if not os.path.exists(path):
            from . import build_oracle
            build_oracle.build()
        _lib = C.CDLL(path)
        _lib.bfe_oracle_threads.restype = C.c_int
    return _lib


def threads():
    return int(load().bfe_oracle_threads())


def use_all_cores():
    """Run the OpenMP loops on every core this process may use, whatever OMP_NUM_THREADS says (torchrun sets it
    to 1 for its workers).  Returns the thread count now in force."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    load().bfe_oracle_set_threads(int(n))
    return threads()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f8(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def eof_geom(g):
    return EofGeom(g['mmax'], g['norder'], g['numx'], g['numy'], g['cmap'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                   g['ascale'], g['hscale'])


def eof_accumulate(x, y, z, m, potC, potS, g):
    lib = load()
    x, y, z, m, potC, potS = (_f8(a) for a in (x, y, z, m, potC, potS))
    c = np.zeros((g['mmax'] + 1, g['norder'])); s = np.zeros_like(c)
    gg = eof_geom(g)
    lib.bfe_oracle_eof_accumulate(C.byref(gg), _p(potC), _p(potS), C.c_long(x.size), _p(x), _p(y), _p(z), _p(m), _p(c), _p(s))
    return c, s


def eof_force(x, y, z, cosc, sinc, T, g, m1=0, m2=1000):
    lib = load()
    x, y, z, cosc, sinc = (_f8(a) for a in (x, y, z, cosc, sinc))
    tabs = [_f8(T[k]) for k in ('potC', 'rforceC', 'zforceC', 'potS', 'rforceS', 'zforceS')]
    arr = (C.c_void_p * 6)(*[t.ctypes.data for t in tabs])
    out = np.zeros((6, x.size))
    gg = eof_geom(g)
    lib.bfe_oracle_eof_force(C.byref(gg), arr, _p(cosc), _p(sinc), C.c_int(m1), C.c_int(m2), C.c_long(x.size), _p(x), _p(y),
                             _p(z), *[_p(out[i]) for i in range(6)])
    return out


def factorial_return(lmax):
    f = np.zeros((lmax + 1, lmax + 1))
    for l in range(lmax + 1):
        for m in range(l + 1):
            f[l, m] = np.sqrt((0.5 * l + 0.25) / np.pi * np.exp(gammaln(1.0 + l - m) - gammaln(1.0 + l + m)))
            if m != 0:
                f[l, m] *= np.sqrt(2.)
    return f


def sl_accumulate(x, y, z, m, lmax, nmax, ev, ef, xi, p0, cmap, scale, no_odd=False):
    lib = load()
    x, y, z, m, ev, ef, xi, p0 = (_f8(a) for a in (x, y, z, m, ev, ef, xi, p0))
    fac = _f8(factorial_return(lmax))
    out = np.zeros(((lmax + 1) ** 2, nmax))
    gg = SlGeom(lmax, nmax, xi.size, cmap, scale)
    lib.bfe_oracle_sl_accumulate(C.byref(gg), _p(ev), _p(ef), _p(xi), _p(p0), _p(fac), C.c_int(int(no_odd)), C.c_long(x.size),
                                 _p(x), _p(y), _p(z), _p(m), _p(out))
    return out
