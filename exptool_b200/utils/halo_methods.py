"""
halo_methods.py -- SL cache / spherical model table ingest
(exptool/utils/halo_methods.py:56-220).  Host side, runs once per basis.
"""
import numpy as np
from scipy import interpolate

from ..basis.compatibility import xi_to_r


def read_sph_model_table(modelfile):
    '''halo_methods.py:56-62: text, '!' comments, five header lines skipped, columns R D M P.
    (The reference also prints row 0; we do not.)'''
    A = np.genfromtxt(modelfile, comments='!', skip_header=5)
    return A[:, 0], A[:, 1], A[:, 2], A[:, 3]


def parse_slgrid(file, verbose=0):
    '''halo_methods.py:65-91'''
    with open(file, 'rb') as f:
        a = np.fromfile(f, dtype=np.uint32, count=4)
        b = np.fromfile(f, dtype=np.float64, count=3)
    return int(a[0]), int(a[1]), int(a[2]), int(a[3]), b[0], b[1], b[2]


def read_cached_table(file, verbose=0, retall=True):
    '''
    halo_methods.py:96-169.  Layout: u4[4] lmax,nmax,numr,cmap; f8[3] rmin,rmax,scale;
    per l: u4 l, f8[nmax] evtable[l], nmax rows f8[numr] eftable[l,n].
    '''
    with open(file, 'rb') as f:
        a = np.fromfile(f, dtype=np.uint32, count=4)
        lmax, nmax, numr, cmap = int(a[0]), int(a[1]), int(a[2]), int(a[3])
        b = np.fromfile(f, dtype='<f8', count=3)
        rmin, rmax, scale = b[0], b[1], b[2]
        ltable = np.zeros(lmax + 1)
        evtable = np.ones([lmax + 1, nmax])
        eftable = np.ones([lmax + 1, nmax, numr])
        for l in range(0, lmax + 1):
            ltable[l] = np.fromfile(f, dtype=np.uint32, count=1)[0]
            evtable[l, 0:nmax] = np.fromfile(f, dtype='f8', count=nmax)
            eftable[l] = np.fromfile(f, dtype='f8', count=nmax * numr).reshape(nmax, numr)
    if retall:
        return lmax, nmax, numr, cmap, rmin, rmax, scale, ltable, evtable, eftable
    return ltable, evtable, eftable


def init_table(modelfile, numr, rmin, rmax, cmap=0, scale=1.0, spline=True):
    '''
    halo_methods.py:178-220: put the model potential / 4 pi density on the cache's
    uniform xi grid with SciPy's own cubic splrep/splev (as the reference does).
    '''
    R1, D1, M1, P1 = read_sph_model_table(modelfile)
    fac0 = 4. * np.pi
    if cmap == 1:
        xmin = (rmin / scale - 1.0) / (rmin / scale + 1.0)
        xmax = (rmax / scale - 1.0) / (rmax / scale + 1.0)
    elif cmap == 0:
        xmin = rmin
        xmax = rmax
    else:
        raise ValueError('halo_methods.init_table: cmap=2 is unusable in the reference '
                         '(undefined `log`, halo_methods.py:195-196)')
    dxi = (xmax - xmin) / (numr - 1)
    xi = np.zeros(numr)
    for i in range(0, numr):
        xi[i] = xmin + dxi * i
    r = xi_to_r(xi, cmap, scale)
    if spline:
        pfunc = interpolate.splrep(R1, P1, s=0)
        dfunc = interpolate.splrep(R1, fac0 * D1, s=0)
        p0 = interpolate.splev(r, pfunc, der=0)
        d0 = interpolate.splev(r, dfunc, der=0)
    else:
        idx = np.abs(r[:, None] - R1[None, :]).argmin(axis=1)
        p0 = P1[idx]
        d0 = fac0 * D1[idx]
    return xi, r, p0, d0
