"""
spheresl.py -- drop-in for the hot-path entry points of exptool/basis/spheresl.py.

Same names, argument order and return values as the reference; the arithmetic
runs in libbfe.so (exptool_b200.ops -> include/bfe.h), including the density
outputs den0, den1 of all_eval / all_eval_particles with each function's own
conventions.

Not mirrored: wake grids, EXP-coefficient readers, rotation helpers
(spheresl.py:1501-1764).
"""
import time
from collections import OrderedDict

import numpy as np
from scipy.special import gammaln

from ..utils import halo_methods
from ..io import particle
from .. import ops


# ---------------------------------------------------------------------------
# small host helpers kept for API compatibility
# ---------------------------------------------------------------------------
def factorial_return(lmax):
    '''spheresl.factorial_return (spheresl.py:823-863); the kernels use the same table,
    built in bfe_sl_create with lgamma.'''
    factorial = np.zeros([lmax + 1, lmax + 1])
    for l in range(0, lmax + 1):
        for m in range(0, l + 1):
            factorial[l][m] = np.sqrt((0.5 * l + 0.25) / np.pi * np.exp(gammaln(1.0 + l - m) - gammaln(1.0 + l + m)))
            if m != 0:
                factorial[l][m] *= np.sqrt(2.)
    return factorial


# ---------------------------------------------------------------------------
# device table cache (the reference re-reads the cache + model files in every worker
# and in every all_eval_particles call: spheresl.py:587-588, 1247-1248)
# ---------------------------------------------------------------------------
_FILE_CACHE = OrderedDict()
_ARRAY_CACHE = OrderedDict()
_CACHE_MAX = 4


def _dev_id():
    return ops.torch.cuda.current_device() if ops.torch.cuda.is_available() else -1


def device_tables_from_files(sph_file, model_file):
    """(ops.SLTables, host table dict) for a cache file + model file pair."""
    import os
    key = (os.path.abspath(sph_file), os.path.getmtime(sph_file), os.path.abspath(model_file),
           os.path.getmtime(model_file), _dev_id())
    hit = _FILE_CACHE.get(key)
    if hit is None:
        lmax, nmax, numr, cmap, rmin, rmax, scale, ltable, evtable, eftable = halo_methods.read_cached_table(sph_file)
        xi, r, p0, d0 = halo_methods.init_table(model_file, numr, rmin, rmax, cmap, scale)
        H = ops.SLTables(lmax, nmax, numr, cmap, scale, evtable, eftable, xi, p0, d0)
        hit = (H, dict(lmax=lmax, nmax=nmax, numr=numr, cmap=cmap, rmin=rmin, rmax=rmax, scale=scale,
                       evtable=evtable, eftable=eftable, xi=xi, p0=p0, d0=d0))
        _FILE_CACHE[key] = hit
        while len(_FILE_CACHE) > _CACHE_MAX:
            _FILE_CACHE.popitem(last=False)
    return hit


def _fingerprint(a):
    """as eof._fingerprint: one cache policy for both bases (eof.set_table_cache_mode / eof.invalidate_tables)"""
    from . import eof as _eof
    return _eof._fingerprint(a)


def device_tables(xi, p0, d0, cmap, scale, evtable, eftable):
    """ops.SLTables for host arrays (full table extent; lmax/nmax truncation happens at contraction)."""
    evtable = np.asarray(evtable); eftable = np.asarray(eftable)
    key = (_fingerprint(xi), _fingerprint(p0), _fingerprint(d0), _fingerprint(evtable), _fingerprint(eftable), int(cmap), float(scale),
           _dev_id())
    from . import eof as _eof
    H = _ARRAY_CACHE.get(key) if _eof._CACHE_MODE['mode'] != 'off' else None
    if H is None:
        lmax = evtable.shape[0] - 1
        nmax = evtable.shape[1]
        H = ops.SLTables(lmax, nmax, eftable.shape[2], cmap, scale, evtable, eftable, xi, p0, d0)
        _ARRAY_CACHE[key] = H
        while len(_ARRAY_CACHE) > _CACHE_MAX:
            _ARRAY_CACHE.popitem(last=False)
    return H


def clear_table_cache():
    _FILE_CACHE.clear()
    _ARRAY_CACHE.clear()


# ---------------------------------------------------------------------------
# per-point building blocks -- spheresl.py:106-160, 301-335, 664-770 (device evaluations)
# ---------------------------------------------------------------------------
def _radial(x, lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale, dens, force, pot):
    scalar = np.ndim(x) == 0
    r1 = np.atleast_1d(np.asarray(x, dtype=np.float64))
    H = device_tables(xi, p0, d0, cmap, scale, np.asarray(evtable)[:lmax + 1, :nmax], np.asarray(eftable)[:lmax + 1, :nmax])
    outs = H.radial_matrices(r1, dens=dens, force=force, pot=pot)
    res = []
    for o in outs:
        if o is None:
            continue
        a = ops.to_host(o)
        res.append(a[:, :, 0].copy() if scalar else a)
    return res


def get_halo_dens_pot_force(x, lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale):
    '''spheresl.get_halo_dens_pot_force (spheresl.py:106-160): dens, force, pot matrices (lmax+1, nmax) at radius x
    (or (lmax+1, nmax, n) for an array of radii).  Return order dens, force, pot (160).'''
    return tuple(_radial(x, lmax, nmax, evtable, eftable, xi, d0, p0, cmap, scale, True, True, True))


def get_halo_pot_matrix(x_in, lmax, nmax, evtable, eftable, xi, p0, cmap, scale):
    '''spheresl.get_halo_pot_matrix (spheresl.py:301-335)'''
    return _radial(x_in, lmax, nmax, evtable, eftable, xi, None, p0, cmap, scale, False, False, True)[0]


def get_halo_dens(x, lmax, nmax, evtable, eftable, xi, d0, cmap, scale):
    '''spheresl.get_halo_dens (spheresl.py:163-189): the density matrix alone (p0 is not used by it).'''
    return _radial(x, lmax, nmax, evtable, eftable, xi, d0, np.zeros_like(np.asarray(d0, dtype=np.float64)), cmap, scale,
                   True, False, False)[0]


def legendre_R(lmax, x):
    '''spheresl.legendre_R (spheresl.py:664-700): P_l^m(x) as (lmax+1, lmax+1) (or (.., .., n) for an array x)'''
    scalar = np.ndim(x) == 0
    P, _ = ops.legendre_tables(lmax, np.atleast_1d(np.asarray(x, dtype=np.float64)), derivative=False)
    P = ops.to_host(P)
    return P[:, :, 0].copy() if scalar else P


def dlegendre_R(lmax, x):
    '''spheresl.dlegendre_R (spheresl.py:706-770): (P, dP)'''
    scalar = np.ndim(x) == 0
    P, dP = ops.legendre_tables(lmax, np.atleast_1d(np.asarray(x, dtype=np.float64)), derivative=True)
    P, dP = ops.to_host(P), ops.to_host(dP)
    if scalar:
        return P[:, :, 0].copy(), dP[:, :, 0].copy()
    return P, dP


# ---------------------------------------------------------------------------
# accumulation -- spheresl.py:567-656, 439-475
# ---------------------------------------------------------------------------
def compute_coefficients_solitary(ParticleInstance, sph_file, model_file, verbose=0, no_odd=False):
    '''
    spheresl.compute_coefficients_solitary (spheresl.py:567-656)
    -> expcoef, ((lmax+1)^2, nmax) float64; row l^2 is m=0, then (cos, sin) pairs.
    '''
    H, _ = device_tables_from_files(sph_file, model_file)
    x, y, z, m = particle.particle_arrays(ParticleInstance)
    return H.accumulate(x, y, z, m, no_odd=no_odd).cpu().numpy()


class SL_Object(object):
    '''spheresl.SL_Object (spheresl.py:1377-1386)'''
    time = None
    dump = None
    comp = None
    nbodies = None
    lmax = None
    nmax = None
    sph_file = None
    model_file = None
    expcoef = None


def compute_coefficients(PSPInput, sph_file, mod_file, verbose=1, no_odd=False):
    '''
    spheresl.compute_coefficients (spheresl.py:439-475) -> SL_Object.  The reference fans out
    over multiprocessing.Pool and sums on the parent (the sum has dtype=object there); here the
    particles go to the GPU (rank-local under torch.distributed; sharded over the ranks with one
    allreduce inside `with exptool_b200.parallel.sharded():`) and expcoef is a float64 array.
    '''
    from .. import parallel
    SL_Out = SL_Object()
    SL_Out.time = getattr(PSPInput, 'time', None)
    SL_Out.filename = getattr(PSPInput, 'filename', None)
    SL_Out.comp = getattr(PSPInput, 'comp', None)
    x, y, z, m = particle.particle_arrays(PSPInput)
    SL_Out.nbodies = int(m.numel()) if isinstance(m, ops.torch.Tensor) else np.asarray(m).size
    SL_Out.sph_file = sph_file
    SL_Out.model_file = mod_file
    H, T = device_tables_from_files(sph_file, mod_file)
    SL_Out.lmax = T['lmax']
    SL_Out.nmax = T['nmax']
    t1 = time.time()
    if parallel.sharded_api():          # opt-in: `with parallel.sharded():` -- every rank holds the same particle set
        SL_Out.expcoef = parallel.raise_if_poisoned(parallel.sl_accumulate_host(H, x, y, z, m, no_odd=no_odd))
    else:                               # rank-local, like compute_coefficients_solitary
        SL_Out.expcoef = H.accumulate_host(x, y, z, m, no_odd=no_odd)
    if verbose > 0:
        dt = time.time() - t1
        print('spheresl.compute_coefficients: accumulation took {0:3.2f} seconds, or {1:4.2f} microseconds per orbit.'
              .format(dt, 1.e6 * dt / max(SL_Out.nbodies, 1)))
    return SL_Out


# ---------------------------------------------------------------------------
# field evaluation -- spheresl.py:987-1102, 1107-1234, 1240-1362
# ---------------------------------------------------------------------------
def _points(r, costh, phi):
    scalar = np.ndim(r) == 0
    return (scalar, np.atleast_1d(np.asarray(r, dtype=np.float64)), np.atleast_1d(np.asarray(costh, dtype=np.float64)),
            np.atleast_1d(np.asarray(phi, dtype=np.float64)))


def force_eval(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax, evtable, eftable, no_odd=False, verbose=0):
    '''
    spheresl.force_eval (spheresl.py:1107-1234): potr, pott, potp, pot1, pot0.
    lmax / nmax truncate the expansion (1138-1148).  As in the reference, the azimuthal
    factors are cos/sin(l phi) (1173, 1222-1225), not cos/sin(m phi).
    '''
    scalar, r1, c1, p1 = _points(r, costh, phi)
    H = device_tables(xi, p0, d0, cmap, scale, evtable, eftable)
    H.contract(expcoef, l1=0, l2=lmax, nuse=nmax, no_odd=no_odd)
    out = ops.to_host(H.force_eval_points(r1, c1, p1, trig_index_l=True))
    if scalar:
        return tuple(np.float64(v[0]) for v in out)
    return tuple(out)


def all_eval(r, costh, phi, expcoef, xi, p0, d0, cmap, scale, lmax, nmax, evtable, eftable, no_odd=False, verbose=0):
    '''
    spheresl.all_eval (spheresl.py:987-1102): den0, den1, pot0, pot1, potr, pott, potp with
    cos/sin(m phi).  As in the reference den1 is the TOTAL density (monopole included, 1046)
    and both densities carry densfac = 0.25/pi (1092).
    '''
    scalar, r1, c1, p1 = _points(r, costh, phi)
    H = device_tables(xi, p0, d0, cmap, scale, evtable, eftable)
    H.contract(expcoef, l1=0, l2=lmax, nuse=nmax, no_odd=no_odd)
    potr, pott, potp, pot1, pot0 = ops.to_host(H.force_eval_points(r1, c1, p1, trig_index_l=False))
    H.contract_density(expcoef, l1=0, l2=lmax, nuse=nmax, no_odd=no_odd)
    den0, den1 = ops.to_host(H.density_eval_points(r1, c1, p1))
    out = (den0, den1, pot0, pot1, potr, pott, potp)
    if scalar:
        return tuple(np.float64(v[0]) for v in out)
    return out


def all_eval_particles(Particles, expcoef, sph_file, mod_file, verbose, L1=-1000, L2=1000, NO_ODD=False):
    '''
    spheresl.all_eval_particles (spheresl.py:1240-1362):
    den0, den1, pot0, pot1, potr, pott, potp, rr per particle.  The density outputs follow the
    reference line by line: den1 excludes the monopole (1271), its m=0 terms are weighted by
    legs[1][0] (1323), and densfac = 0.25*pi (1351).
    '''
    H, _ = device_tables_from_files(sph_file, mod_file)
    x, y, z, _m = particle.particle_arrays(Particles)
    H.contract(expcoef, l1=L1, l2=L2, no_odd=NO_ODD)
    H.contract_density(expcoef, l1=L1, l2=L2, no_odd=NO_ODD)
    # ONE upload of x, y, z serves the potential and the density evaluation, ONE copy out of the eight rows
    x, y, z = ops.dev(x), ops.dev(y), ops.dev(z)
    out = ops.to_host(ops.torch.cat([H.density(x, y, z), H.force(x, y, z)]))
    den0, den1, pot0, pot1, potr, pott, potp, rr = out
    return den0, den1, pot0, pot1, potr, pott, potp, rr


def eval_particles(ParticleInstance, expcoef, sph_file, mod_file, nprocs=-1, l1=0, l2=1000, verbose=1, no_odd=False):
    '''spheresl.eval_particles (spheresl.py:502-528); `nprocs` accepted for call compatibility.'''
    return all_eval_particles(ParticleInstance, expcoef, sph_file, mod_file, verbose, L1=l1, L2=l2, NO_ODD=no_odd)


# ---------------------------------------------------------------------------
# coefficient dump files -- spheresl.py:1390-1488 (byte layout of SURVEY.md App. B.4)
# ---------------------------------------------------------------------------
def sl_coefficients_to_file(f, SL_Object):
    '''spheresl.py:1390-1410: 324-byte header then expcoef as f8'''
    np.array([SL_Object.time], dtype='f4').tofile(f)
    np.array([SL_Object.filename], dtype='S100').tofile(f)
    np.array([SL_Object.comp], dtype='S8').tofile(f)
    np.array([SL_Object.nbodies], dtype='i4').tofile(f)
    np.array([SL_Object.sph_file], dtype='S100').tofile(f)
    np.array([SL_Object.model_file], dtype='S100').tofile(f)
    np.array([SL_Object.lmax, SL_Object.nmax], dtype='i4').tofile(f)
    np.array(np.asarray(SL_Object.expcoef, dtype=np.float64).reshape(-1, ), dtype='f8').tofile(f)


def save_sl_coefficients(outfile, SL_Object, verbose=0):
    '''spheresl.py:1412-1439 (the reference passes an array to f.seek and raises on NumPy 2;
    the offset formula is kept)'''
    try:
        f = open(outfile, 'rb+')
        f.close()
    except IOError:
        f = open(outfile, 'wb')
        np.array([0], dtype='i4').tofile(f)
        f.close()
    with open(outfile, 'rb+') as f:
        ndumps = int(np.fromfile(f, dtype='i4', count=1)[0]) + 1
        f.seek(0)
        np.array([ndumps], dtype='i4').tofile(f)
        if verbose:
            print('spheresl.save_sl_coefficients: coefficient file currently has {0:d} dumps.'.format(ndumps))
        f.seek(4 + (ndumps - 1) * (8 * ((SL_Object.lmax + 1) * (SL_Object.lmax + 1)) * (SL_Object.nmax) + 324))
        sl_coefficients_to_file(f, SL_Object)


def extract_sl_coefficients(f):
    '''spheresl.py:1467-1488'''
    SL_Obj = SL_Object()
    [SL_Obj.time] = np.fromfile(f, dtype='f4', count=1)
    [SL_Obj.filename] = np.fromfile(f, dtype='S100', count=1)
    [SL_Obj.comp] = np.fromfile(f, dtype='S8', count=1)
    [SL_Obj.nbodies] = np.fromfile(f, dtype='i4', count=1)
    [SL_Obj.sph_file] = np.fromfile(f, dtype='S100', count=1)
    [SL_Obj.model_file] = np.fromfile(f, dtype='S100', count=1)
    [SL_Obj.lmax, SL_Obj.nmax] = np.fromfile(f, dtype='i4', count=2)
    flat = np.fromfile(f, dtype='f8', count=(SL_Obj.lmax + 1) * (SL_Obj.lmax + 1) * SL_Obj.nmax)
    SL_Obj.expcoef = flat.reshape([(SL_Obj.lmax + 1) * (SL_Obj.lmax + 1), SL_Obj.nmax])
    return SL_Obj


def restore_sl_coefficients(infile):
    '''spheresl.py:1442-1464: (last SL_Object, dict np.round(time, 3) -> SL_Object)'''
    SL_Dict = OrderedDict()
    SL_Out = None
    with open(infile, 'rb') as f:
        [ndumps] = np.fromfile(f, dtype='i4', count=1)
        f.seek(4)
        for step in range(0, ndumps):
            SL_Out = extract_sl_coefficients(f)            # a short / corrupt dump raises, as in the reference
            SL_Dict[np.round(SL_Out.time, 3)] = SL_Out       # keyed by the ROUNDED time (spheresl.py:1461)
    return SL_Out, SL_Dict
