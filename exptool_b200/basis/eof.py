"""
eof.py -- drop-in for the hot-path entry points of exptool/basis/eof.py.

Same names, argument order, return shapes and NumPy return types as the
reference; the arithmetic runs in libbfe.so on the current CUDA device
(exptool_b200.ops -> include/bfe.h).  Table readers and coefficient-file I/O
are host-side NumPy, as in the reference.

Not mirrored (outside the path): wake grids, plotting.
"""
import os
import time
from collections import OrderedDict

import numpy as np

from . import compatibility
from ..io import particle
from .. import ops

r_to_xi = compatibility.r_to_xi
z_to_y = compatibility.z_to_y


# ---------------------------------------------------------------------------
# cache-file readers (host) -- eof.py:98-313
# ---------------------------------------------------------------------------
def _read_header(f):
    tmagic = np.fromfile(f, dtype='<i4', count=1)
    hmagic = 0xc0a57a1
    if tmagic.size and tmagic[0] == hmagic:
        # new-style header (eof.py:126-157): magic, length, YAML block
        import yaml
        ssize = int(np.fromfile(f, dtype='<i4', count=1)[0])
        data = yaml.safe_load(f.read(ssize).decode('utf-8').rstrip('\x00'))
        cmap = data['cmap'] if 'cmap' in data else data['cmapr']
        hdr = dict(mmax=int(data['mmax']), numx=int(data['numx']), numy=int(data['numy']), nmax=int(data['nmax']),
                   norder=int(data['norder']), dens=int(data['dens']), cmap=int(cmap), rmin=float(data['rmin']),
                   rmax=float(data['rmax']), ascale=float(data['ascl']), hscale=float(data['hscl']),
                   cylmass=float(data['cmass']), time=float(data['time']))
        offset = 8 + ssize
    else:
        # old-style header (eof.py:159-181): 7 x u4 + 6 x f8 = 76 bytes
        f.seek(0, 0)
        a = np.fromfile(f, dtype=np.uint32, count=7)
        b = np.fromfile(f, dtype='<f8', count=6)
        hdr = dict(mmax=int(a[0]), numx=int(a[1]), numy=int(a[2]), nmax=int(a[3]), norder=int(a[4]),
                   dens=int(a[5]), cmap=int(a[6]), rmin=float(b[0]), rmax=float(b[1]), ascale=float(b[2]),
                   hscale=float(b[3]), cylmass=float(b[4]), time=float(b[5]))
        offset = 76
    return hdr, offset


def eof_params(file, verbose=0):
    '''eof.eof_params (eof.py:98-200): rmin,rmax,numx,numy,mmax,norder,ascale,hscale,cmap,dens'''
    with open(file, 'rb') as f:
        h, _ = _read_header(f)
    if verbose:
        print('eof.eof_params: The parameters for this EOF file are:')
        print('RMIN={0:5.4f},RMAX={1:5.4f}'.format(h['rmin'], h['rmax']))
        print('MMAX={0:d}'.format(h['mmax']))
        print('NORDER={0:d}'.format(h['norder']))
        print('NMAX={0:d}'.format(h['nmax']))
        print('NUMX,NUMY={0:d},{1:d}'.format(h['numx'], h['numy']))
        print('DENS,CMAP={0:d},{1:d}'.format(h['dens'], h['cmap']))
        print('ASCALE,HSCALE={0:5.4f},{1:5.4f}'.format(h['ascale'], h['hscale']))
        print('CYLMASS={0:5.4f}'.format(h['cylmass']))
        print('TNOW={0:5.4f}'.format(h['time']))
    return h['rmin'], h['rmax'], h['numx'], h['numy'], h['mmax'], h['norder'], h['ascale'], h['hscale'], h['cmap'], h['dens']


def parse_eof(file, verbose=0):
    '''
    eof.parse_eof (eof.py:224-313): potC,rforceC,zforceC,densC,potS,rforceS,zforceS,densS,
    each (mmax+1, norder, numx+1, numy+1); sine arrays have a zero m=0 plane.
    One bulk read per block instead of the reference's row-by-row np.fromfile loop.
    '''
    with open(file, 'rb') as f:
        h, offset = _read_header(f)
        f.seek(offset)
        M, N, NX, NY = h['mmax'] + 1, h['norder'], h['numx'] + 1, h['numy'] + 1
        nf = 4 if h['dens'] == 1 else 3
        cosb = np.fromfile(f, dtype='<f8', count=M * N * nf * NX * NY).reshape(M, N, nf, NX, NY)
        sinb = np.fromfile(f, dtype='<f8', count=(M - 1) * N * nf * NX * NY).reshape(M - 1, N, nf, NX, NY)
    shape = (M, N, NX, NY)
    out = []
    for blk, first in ((cosb, 0), (sinb, 1)):
        for k in range(4):
            a = np.zeros(shape)
            if k < nf:
                a[first:] = blk[:, :, k]
            out.append(a)
    potC, rforcec, zforcec, densc, potS, rforces, zforces, denss = out
    return potC, rforcec, zforcec, densc, potS, rforces, zforces, denss


def read_eof_file(file):
    """dictionary wrapper for parse_eof (eof.py:205-221)"""
    keys = ('potC', 'rforceC', 'zforceC', 'densC', 'potS', 'rforceS', 'zforceS', 'densS')
    return dict(zip(keys, parse_eof(file)))


def set_table_params(RMAX=20.0, RMIN=0.001, ASCALE=0.01, HSCALE=0.001, NUMX=128, NUMY=64, CMAP=0):
    '''eof.set_table_params (eof.py:316-347): XMIN,XMAX,dX,YMIN,YMAX,dY'''
    M_SQRT1_2 = np.sqrt(0.5)
    Rtable = M_SQRT1_2 * RMAX
    XMIN = r_to_xi(RMIN * ASCALE, CMAP, ASCALE)
    XMAX = r_to_xi(Rtable * ASCALE, CMAP, ASCALE)
    dX = (XMAX - XMIN) / NUMX
    YMIN = z_to_y(-Rtable * ASCALE, HSCALE)
    YMAX = z_to_y(Rtable * ASCALE, HSCALE)
    dY = (YMAX - YMIN) / NUMY
    return XMIN, XMAX, dX, YMIN, YMAX, dY


# ---------------------------------------------------------------------------
# device table cache: the reference API passes the NumPy tables on every call
# ---------------------------------------------------------------------------
_TABLE_CACHE = OrderedDict()
_TABLE_CACHE_MAX = 4


# How a host table is recognised on the next call (the reference API passes the NumPy tables on EVERY call):
#   'sampled' (default): address, shape, dtype + the sum of 509 evenly spaced values and of the first / last 64 -- an
#                        in-place edit that misses every sampled value goes unnoticed: call invalidate_tables() after
#                        editing a table in place (or use 'full');
#   'full'             : address, shape, dtype + a hash of ALL bytes (hashlib.blake2b; ~25 ms per 17 MB table per call);
#   'off'              : no cache, tables are uploaded on every call.
_CACHE_MODE = {'mode': os.environ.get('BFE_TABLE_CACHE', 'sampled')}


def set_table_cache_mode(mode):
    if mode not in ('sampled', 'full', 'off'):
        raise ValueError("table cache mode must be 'sampled', 'full' or 'off'")
    _CACHE_MODE['mode'] = mode
    clear_table_cache()


def invalidate_tables(*arrays):
    """Forget the device copies made from these host arrays (all cached tables if none are given): call it after
    editing a table in place.  Also drops the SL caches (spheresl.clear_table_cache)."""
    if not arrays:
        clear_table_cache()
    else:
        ptrs = set(np.asarray(a).__array_interface__['data'][0] for a in arrays if a is not None)
        for key in [k for k in _TABLE_CACHE if any(fp is not None and fp[0] in ptrs for fp in k[0])]:
            del _TABLE_CACHE[key]
    from . import spheresl as _sl
    _sl.clear_table_cache()


def _fingerprint(a):
    if a is None:
        return None
    a = np.asarray(a)
    flat = a.reshape(-1)
    head = (a.__array_interface__['data'][0], a.shape, a.dtype.str)
    if _CACHE_MODE['mode'] == 'full':
        import hashlib
        return head + (hashlib.blake2b(np.ascontiguousarray(a).view(np.uint8).reshape(-1), digest_size=16).hexdigest(),)
    step = max(1, flat.size // 509)
    return head + (float(flat[::step].sum()), float(flat[:64].sum()), float(flat[-64:].sum()))


def device_tables(potC, potS, MMAX, NMAX, XMIN, dX, YMIN, dY, NUMX, NUMY, ASCALE, HSCALE, CMAP,
                  rforceC=None, zforceC=None, rforceS=None, zforceS=None):
    """ops.EOFTables for these host tables (uploaded once, then cached)."""
    key = (tuple(_fingerprint(t) for t in (potC, potS, rforceC, zforceC, rforceS, zforceS)),
           int(MMAX), int(NMAX), float(XMIN), float(dX), float(YMIN), float(dY), int(NUMX), int(NUMY),
           float(ASCALE), float(HSCALE), int(CMAP), ops.torch.cuda.current_device() if ops.torch.cuda.is_available() else -1)
    E = _TABLE_CACHE.get(key) if _CACHE_MODE['mode'] != 'off' else None
    if E is None:
        E = ops.EOFTables(potC, potS, MMAX, NMAX, XMIN, dX, YMIN, dY, NUMX, NUMY, ASCALE, HSCALE, CMAP,
                          rforceC=rforceC, zforceC=zforceC, rforceS=rforceS, zforceS=zforceS)
        _TABLE_CACHE[key] = E
        while len(_TABLE_CACHE) > _TABLE_CACHE_MAX:
            _TABLE_CACHE.popitem(last=False)
    else:
        _TABLE_CACHE.move_to_end(key)
    return E


def clear_table_cache():
    _TABLE_CACHE.clear()


# ---------------------------------------------------------------------------
# per-point building blocks -- eof.py:354-457 (evaluated on the device like everything else)
# ---------------------------------------------------------------------------
def return_bins(r, z, rmin=0, dR=0, zmin=0, dZ=0, numx=0, numy=0, ASCALE=0.01, HSCALE=0.001, CMAP=0):
    '''
    eof.return_bins (eof.py:354-427): X, Y (exact table coordinates), ix, iy (truncated bins).
    The lower edge clamps ix and X; the upper edge clamps ix only, so X extrapolates (414-415).
    Scalars in give 0-d arrays out, as in the reference.
    '''
    scalar = np.ndim(r) == 0
    r1 = np.atleast_1d(np.asarray(r, dtype=np.float64))
    z1 = np.atleast_1d(np.asarray(z, dtype=np.float64))
    X, Y, ix, iy = ops.eof_return_bins(r1, z1, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP)
    X, Y = ops.to_host(X), ops.to_host(Y)
    ix, iy = ops.to_host(ix).astype(int), ops.to_host(iy).astype(int)
    if scalar:
        return np.squeeze(X), np.squeeze(Y), np.squeeze(ix), np.squeeze(iy)
    return X, Y, ix, iy


def get_pot(r, z, cos_array, sin_array, rmin=0, dR=0, zmin=0, dZ=0, numx=0, numy=0, fac=1.0, MMAX=6, NMAX=18,
            ASCALE=0.01, HSCALE=0.001, CMAP=0):
    '''
    eof.get_pot (eof.py:430-457): bilinear interpolation of every (m, n) cosine and sine table at the
    points -> Vc, Vs of shape (m, n, npoints).  The m=0 sine plane is taken as zero (parse_eof, eof.py:293).
    '''
    cos_array = np.asarray(cos_array); sin_array = np.asarray(sin_array)
    r1 = np.atleast_1d(np.asarray(r, dtype=np.float64))
    z1 = np.atleast_1d(np.asarray(z, dtype=np.float64))
    E = device_tables(cos_array, sin_array, cos_array.shape[0] - 1, cos_array.shape[1], rmin, dR, zmin, dZ, numx, numy,
                      ASCALE, HSCALE, CMAP)
    Vc, Vs = E.get_pot(r1, z1, fac=fac)
    return ops.to_host(Vc), ops.to_host(Vs)


# ---------------------------------------------------------------------------
# accumulation -- eof.py:492-640, 1415-1455, 1158-1260
# ---------------------------------------------------------------------------
def accumulate(ParticleInstance, potC, potS, MMAX, NMAX, XMIN, dX, YMIN, dY, NUMX, NUMY, ASCALE, HSCALE, CMAP,
               verbose=0, no_odd=False, VAR=0):
    '''
    eof.accumulate (eof.py:492-640) -> accum_cos, accum_sin, each (MMAX+1, NMAX) float64.
    `no_odd` is accepted and, as in the reference (mask computed at 545-548 but never
    applied), has no effect.  VAR > 0 adds the jackknife partitions (554-574): the return is then
    accum_cos, accum_sin, accum_cos2, accum_sin2 (see _accumulate_with_variance).
    '''
    x, y, z, m = particle.particle_arrays(ParticleInstance)
    E = device_tables(potC, potS, MMAX, NMAX, XMIN, dX, YMIN, dY, NUMX, NUMY, ASCALE, HSCALE, CMAP)
    if not VAR:
        return E.accumulate_host(x, y, z, m)
    return _accumulate_with_variance(E, x, y, z, m, int(VAR), MMAX, NMAX)


def _accumulate_with_variance(E, x, y, z, m, nvar, MMAX, NMAX):
    '''
    The VAR branch of eof.accumulate (eof.py:554-574 / 617-637): besides the coefficients, `VAR` jackknife
    partitions, each the coefficient sum over floor(sqrt(N)) particles drawn WITH replacement by
    np.random.randint from NumPy's global generator -- the same calls in the same order as the reference, so a
    caller who seeds np.random gets the reference's partitions.  Returns accum_cos, accum_sin,
    accum_cos2 (VAR, M+1, N), accum_sin2 (VAR, M+1, N).
    '''
    xd, yd, zd, md = ops.dev(x), ops.dev(y), ops.dev(z), ops.dev(m)
    c, s = E.accumulate(xd, yd, zd, md)
    n = int(xd.numel())
    cos2 = np.zeros([nvar, MMAX + 1, NMAX]); sin2 = np.zeros([nvar, MMAX + 1, NMAX])
    parts = []
    for T in range(nvar):
        use = np.random.randint(n, size=int(np.floor(np.sqrt(n))))                  # eof.py:567
        idx = ops.torch.from_numpy(use).to(xd.device)
        parts.append(ops.torch.stack(E.accumulate(xd[idx], yd[idx], zd[idx], md[idx])))
    allp = ops.to_host(ops.torch.stack(parts)) if parts else np.zeros((0, 2, MMAX + 1, NMAX))
    for T in range(nvar):
        cos2[T] = allp[T, 0]; sin2[T] = allp[T, 1]
    cs = ops.to_host(ops.torch.stack([c, s]))
    return cs[0], cs[1], cos2, sin2


def make_coefficients_multi(ParticleInstance, nprocs, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy,
                            ascale, hscale, cmap, verbose=0, no_odd=False, VAR=False):
    '''
    eof.make_coefficients_multi (eof.py:1415-1455).  The reference splits the particles over
    `nprocs` worker processes and sums the partial coefficients on the parent; here one GPU
    takes the whole set (rank-local under torch.distributed, like eof.accumulate).  Inside
    `with exptool_b200.parallel.sharded():` every rank takes its block of the SAME particle set
    and the partial coefficients are summed with one allreduce over NVLink (exptool_b200.parallel).
    `nprocs` is accepted for call compatibility.
    '''
    if VAR:
        # the reference sums per-worker tuples of unequal shapes here (np.array(a_coeffs), eof.py:1440), which NumPy
        # >= 1.24 rejects; the partitions are drawn over the whole set instead, as in the single-process branch
        return accumulate(ParticleInstance, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale,
                          cmap, verbose=verbose, no_odd=no_odd, VAR=VAR)
    t1 = time.time()
    from .. import parallel
    x, y, z, m = particle.particle_arrays(ParticleInstance)
    E = device_tables(potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap)
    if parallel.sharded_api():
        c, s = parallel.eof_accumulate_host(E, x, y, z, m)
        parallel.raise_if_poisoned(c); parallel.raise_if_poisoned(s)
    else:
        c, s = E.accumulate_host(x, y, z, m)
    if verbose:
        dt = time.time() - t1
        print('eof.make_coefficients_multi: Accumulation took {0:3.2f} seconds, or {1:4.2f} microseconds per orbit.'
              .format(dt, 1.e6 * dt / max(len(x), 1)))
    return c, s


class EOF_Object(object):
    '''eof.EOF_Object (eof.py:1627-1636)'''
    time = None
    dump = None
    comp = None
    nbodies = None
    mmax = None
    nmax = None
    eof_file = None
    cos = None
    sin = None


def compute_coefficients(PSPInput, eof_file, verbose=1, no_odd=False, nprocs_max=-1, VAR=False, nanblock=False):
    '''
    eof.compute_coefficients (eof.py:1158-1260) -> EOF_Object.
    NaN positions are moved to the origin unless nanblock (the reference's intent at
    1190-1204; its own `if nanvals > 0` raises on NumPy >= 2.2).
    '''
    x, y, z, m = particle.particle_arrays(PSPInput)
    on_device = isinstance(x, ops.torch.Tensor) and x.is_cuda
    if on_device:
        # device-resident snapshot (Fields.total_coefficients after the transforms): same NaN rule on the device
        bad = ops.torch.isnan(x) | ops.torch.isnan(y) | ops.torch.isnan(z)
        if bool(bad.any()):
            print('eof.compute_coefficients: NaN values found in output file {}.'.format(getattr(PSPInput, 'filename', None)))
            if not nanblock:
                x, y, z = [ops.torch.where(bad, ops.torch.zeros_like(a), a) for a in (x, y, z)]
    else:
        x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64); z = np.asarray(z, dtype=np.float64)
        nanvals = np.where(np.isnan(x) | np.isnan(y) | np.isnan(z))[0]
        if nanvals.size > 0:
            print('eof.compute_coefficients: NaN values found in output file {}.'.format(getattr(PSPInput, 'filename', None)))
            if not nanblock:
                x = x.copy(); y = y.copy(); z = z.copy()
                x[nanvals] = 0.; y[nanvals] = 0.; z[nanvals] = 0.
    EOF_Out = EOF_Object()
    EOF_Out.time = getattr(PSPInput, 'time', None)
    EOF_Out.filename = getattr(PSPInput, 'filename', None)
    EOF_Out.comp = getattr(PSPInput, 'comp', None)
    EOF_Out.nbodies = int(m.numel()) if isinstance(m, ops.torch.Tensor) else np.asarray(m).size
    EOF_Out.eof_file = eof_file
    potC, rforceC, zforceC, densC, potS, rforceS, zforceS, densS = parse_eof(eof_file)
    rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, dens = eof_params(eof_file, verbose=(verbose > 1))
    XMIN, XMAX, dX, YMIN, YMAX, dY = set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ascale, HSCALE=hscale,
                                                      NUMX=numx, NUMY=numy, CMAP=cmap)
    EOF_Out.mmax = mmax
    EOF_Out.nmax = norder
    res = make_coefficients_multi((x, y, z, m), 1, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy,
                                  ascale, hscale, cmap, verbose=verbose, no_odd=no_odd, VAR=VAR)
    EOF_Out.cos, EOF_Out.sin = res[0], res[1]
    if VAR:                                           # eof.py:1254-1256
        EOF_Out.cos2, EOF_Out.sin2 = res[2], res[3]
    return EOF_Out


# ---------------------------------------------------------------------------
# field evaluation -- eof.py:756-870, 989-1144, 1263-1317
# ---------------------------------------------------------------------------
def _scalar_or_array(vals, scalar):
    if scalar:
        return tuple(np.float64(v[0]) for v in vals)
    return tuple(vals)


def force_eval(r, z, phi, accum_cos, accum_sin, potC, rforceC, zforceC, potS, rforceS, zforceS,
               rmin=0, dR=0, zmin=0, dZ=0, numx=0, numy=0, fac=1.0, MMAX=6, NMAX=18, ASCALE=0.0, HSCALE=0.0,
               CMAP=0, no_odd=False, perturb=False):
    '''
    eof.force_eval (eof.py:756-870): (fr+fr0, fp, fz+fz0, p+p0, p0), or
    (fr, fp, fz, p, p0, fr0, fz0) if perturb.  r, z, phi may be scalars (as the
    reference is called) or equal-length arrays (batched extension).
    '''
    scalar = np.ndim(r) == 0
    r1 = np.atleast_1d(np.asarray(r, dtype=np.float64))
    z1 = np.atleast_1d(np.asarray(z, dtype=np.float64))
    p1 = np.atleast_1d(np.asarray(phi, dtype=np.float64))
    E = device_tables(potC, potS, MMAX, NMAX, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP,
                      rforceC=rforceC, zforceC=zforceC, rforceS=rforceS, zforceS=zforceS)
    if not perturb:
        E.contract(accum_cos, accum_sin, m1=0, m2=MMAX, nuse=NMAX, no_odd=no_odd)
        fr, fp, fz, p, p0 = ops.to_host(E.force_eval_points(r1, z1, p1))
        return _scalar_or_array((fr, fp, fz, p, p0), scalar)
    E.contract(accum_cos, accum_sin, m1=1, m2=MMAX, nuse=NMAX, no_odd=no_odd)
    fr, fp, fz, p, _ = ops.to_host(E.force_eval_points(r1, z1, p1))
    E.contract(accum_cos, accum_sin, m1=0, m2=0, nuse=NMAX, no_odd=False)
    fr0, _, fz0, _, p0 = ops.to_host(E.force_eval_points(r1, z1, p1))
    return _scalar_or_array((fr, fp, fz, p, p0, fr0, fz0), scalar)


def accumulated_eval(r, z, phi, accum_cos, accum_sin, potC, rforceC, zforceC, densC, potS, rforceS, zforceS, densS,
                     rmin=0, dR=0, zmin=0, dZ=0, numx=0, numy=0, fac=1.0, MMAX=6, NMAX=18, ASCALE=0.0, HSCALE=0.0,
                     CMAP=0, no_odd=False):
    '''
    eof.accumulated_eval (eof.py:874-929): p0, p, fr, fp, fz, d0, d at a point (or equal-length arrays of
    points).  Here p and d are the TOTAL sums (m=0 included; p0, d0 are the m=0 parts), and no_odd skips
    odd m entirely (896-897).  Density tables that are scalars / None give d0 = d = 0.
    '''
    scalar = np.ndim(r) == 0
    r1 = np.atleast_1d(np.asarray(r, dtype=np.float64))
    z1 = np.atleast_1d(np.asarray(z, dtype=np.float64))
    p1 = np.atleast_1d(np.asarray(phi, dtype=np.float64))
    E = device_tables(potC, potS, MMAX, NMAX, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP,
                      rforceC=rforceC, zforceC=zforceC, rforceS=rforceS, zforceS=zforceS)
    E.contract(accum_cos, accum_sin, m1=0, m2=MMAX, nuse=NMAX, no_odd=no_odd)
    fr, fp, fz, p, p0 = ops.to_host(E.force_eval_points(r1, z1, p1))
    if np.ndim(densC) == 4 and np.ndim(densS) == 4:
        D = device_tables(densC, densS, MMAX, NMAX, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP,
                          rforceC=densC, zforceC=densC, rforceS=densS, zforceS=densS)
        D.contract(accum_cos, accum_sin, m1=0, m2=MMAX, nuse=NMAX, no_odd=no_odd)
        _, _, _, d, d0 = ops.to_host(D.force_eval_points(r1, z1, p1))
    else:
        d0 = np.zeros_like(p0); d = np.zeros_like(p0)
    return _scalar_or_array((p0, p, fr, fp, fz, d0, d), scalar)


def accumulated_eval_particles(Particles, accum_cos, accum_sin, potC=0, rforceC=0, zforceC=0, potS=0, rforceS=0,
                               zforceS=0, rmin=0, dR=0, zmin=0, dZ=0, numx=0, numy=0, MMAX=6, NMAX=18, ASCALE=0.0,
                               HSCALE=0.0, CMAP=0, m1=0, m2=1000, verbose=1, density=False, eof_file=''):
    '''
    eof.accumulated_eval_particles (eof.py:989-1144): p0, p, fr, fp, fz, R, one value per particle.
    p excludes m=0; fr, fz include it; only m1 <= m <= m2 contribute.
    '''
    dens = 0                                                       # eof.py:1039
    densC = densS = None
    if eof_file != '':
        potC, rforceC, zforceC, densC, potS, rforceS, zforceS, densS = parse_eof(eof_file)
        rmin, rmax, numx, numy, MMAX, NMAX, ASCALE, HSCALE, CMAP, dens = eof_params(eof_file)
        rmin, rmax, dR, zmin, zmax, dZ = set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ASCALE, HSCALE=HSCALE,
                                                          NUMX=numx, NUMY=numy, CMAP=CMAP)
    if density and dens == 0:
        # eof.py:1048-1050 (the reference only resets the flag when verbose > 0 and then fails on the missing
        # density tables; here the flag is always dropped)
        if verbose > 0:
            print('eof.accumulated_eval_particles: cannot compute density (functions not specified). moving on without...')
        density = False
    x, y, z, _ = particle.particle_arrays(Particles)
    E = device_tables(potC, potS, MMAX, NMAX, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP,
                      rforceC=rforceC, zforceC=zforceC, rforceS=rforceS, zforceS=zforceS)
    E.contract(accum_cos, accum_sin, m1=m1, m2=m2)
    p0, p, fr, fp, fz, R = E.force_host(x, y, z)
    if not density:
        return p0, p, fr, fp, fz, R
    # density (eof.py:1106, 1122, 1136-1138): d0, d are the sums that give p0, p with densC / densS in place of
    # potC / potS -- the same contraction and interpolation kernels run on a second table handle built from the
    # density tables (its force tables are unused placeholders)
    D = device_tables(densC, densS, MMAX, NMAX, rmin, dR, zmin, dZ, numx, numy, ASCALE, HSCALE, CMAP,
                      rforceC=densC, zforceC=densC, rforceS=densS, zforceS=densS)
    D.contract(accum_cos, accum_sin, m1=m1, m2=m2)
    d0, d = D.force_host(x, y, z)[:2]
    return p0, p, d0, d, fr, fp, fz, R


def compute_forces(PSPInput, EOF_Object, verbose=1, nprocs=-1, m1=0, m2=1000, density=False):
    '''eof.compute_forces (eof.py:1263-1317): p0, p, fr, fp, fz, r'''
    potC, rforceC, zforceC, densC, potS, rforceS, zforceS, densS = parse_eof(EOF_Object.eof_file)
    rmin, rmax, numx, numy, mmax, norder, ascale, hscale, cmap, dens = eof_params(EOF_Object.eof_file)
    XMIN, XMAX, dX, YMIN, YMAX, dY = set_table_params(RMAX=rmax, RMIN=rmin, ASCALE=ascale, HSCALE=hscale,
                                                      NUMX=numx, NUMY=numy, CMAP=cmap)
    return accumulated_eval_particles(PSPInput, EOF_Object.cos, EOF_Object.sin, potC, rforceC, zforceC, potS,
                                      rforceS, zforceS, rmin=XMIN, dR=dX, zmin=YMIN, dZ=dY, numx=numx, numy=numy,
                                      MMAX=mmax, NMAX=norder, ASCALE=ascale, HSCALE=hscale, CMAP=cmap, m1=m1, m2=m2,
                                      verbose=verbose, density=density)


# ---------------------------------------------------------------------------
# the reference's process fan-out helpers (eof.py:1333-1455, 1521-1590), kept call-compatible.  The reference
# cuts the particle set into `nprocs` blocks for a multiprocessing.Pool; here the blocks are consecutive device
# launches (or ranks, see parallel.py), and the partition is the reference's own.
# ---------------------------------------------------------------------------
def redistribute_particles(ParticleInstance, divisions):
    '''eof.redistribute_particles (eof.py:1333-1362): block 0 takes the remainder.'''
    x, y, z, m = particle.particle_arrays(ParticleInstance)
    from .. import parallel
    holders = []
    for lo, hi in parallel.shard_bounds(len(x), divisions):
        h = particle.holder()
        h.xpos, h.ypos, h.zpos, h.mass = x[lo:hi], y[lo:hi], z[lo:hi], m[lo:hi]
        holders.append(h)
    return holders


def multi_accumulate(holding, nprocs, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap,
                     verbose=0, no_odd=False, VAR=False):
    '''eof.multi_accumulate (eof.py:1371-1412): list of (accum_cos, accum_sin), one pair per block.'''
    return [accumulate(holding[i], potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap,
                       verbose=verbose if i == 0 else 0, no_odd=no_odd, VAR=VAR) for i in range(nprocs)]


def mix_outputs(MultiOutput, density=False):
    '''
    eof.mix_outputs (eof.py:1547-1590): concatenate per-block outputs in block order.  (The reference's
    density branch never advances its write offset, eof.py:1567-1575; the blocks are concatenated here.)
    '''
    ncol = 8 if density else 6
    return tuple(np.concatenate([np.asarray(MultiOutput[i][c], dtype=np.float64) for i in range(len(MultiOutput))])
                 for c in range(ncol))


def find_forces_multi(ParticleInstance, nprocs, a_cos, a_sin, potC, rforceC, zforceC, potS, rforceS, zforceS, XMIN, dX,
                      YMIN, dY, numx, numy, mmax, norder, ascale, hscale, cmap, m1=0, m2=1000, verbose=0, density=False):
    '''
    eof.find_forces_multi (eof.py:1521-1541).  As in the reference the m window passed in is ignored
    (0..1000 is hard-coded at 1528).  One device pass over the whole set replaces the Pool.
    '''
    return accumulated_eval_particles(ParticleInstance, a_cos, a_sin, potC=potC, rforceC=rforceC, zforceC=zforceC,
                                      potS=potS, rforceS=rforceS, zforceS=zforceS, rmin=XMIN, dR=dX, zmin=YMIN, dZ=dY,
                                      numx=numx, numy=numy, MMAX=mmax, NMAX=norder, ASCALE=ascale, HSCALE=hscale,
                                      CMAP=cmap, m1=0, m2=1000, verbose=verbose, density=density)


# ---------------------------------------------------------------------------
# coefficient dump files -- eof.py:1646-1760 (byte-compatible)
# ---------------------------------------------------------------------------
def eof_coefficients_to_file(f, EOF_Object):
    '''eof.py:1646-1667: 224-byte header then cos, sin as f8'''
    np.array([EOF_Object.time], dtype='f4').tofile(f)
    np.array([EOF_Object.filename], dtype='S100').tofile(f)
    np.array([EOF_Object.comp], dtype='S8').tofile(f)
    np.array([EOF_Object.nbodies], dtype='i4').tofile(f)
    np.array([EOF_Object.eof_file], dtype='S100').tofile(f)
    np.array([EOF_Object.mmax, EOF_Object.nmax], dtype='i4').tofile(f)
    np.array(EOF_Object.cos.reshape(-1, ), dtype='f8').tofile(f)
    np.array(EOF_Object.sin.reshape(-1, ), dtype='f8').tofile(f)


def save_eof_coefficients(outfile, EOF_Object, verbose=0):
    '''eof.py:1671-1706: append one dump, bump the leading i4 counter'''
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
            print('eof.save_eof_coefficients: coefficient file currently has {0:d} dumps.'.format(ndumps))
        f.seek(4 + (ndumps - 1) * (16 * (EOF_Object.mmax + 1) * (EOF_Object.nmax) + 224))
        eof_coefficients_to_file(f, EOF_Object)


def extract_eof_coefficients(f):
    '''eof.py:1740-1760'''
    EOF_Obj = EOF_Object()
    [EOF_Obj.time] = np.fromfile(f, dtype='f4', count=1)
    [EOF_Obj.filename] = np.fromfile(f, dtype='S100', count=1)
    [EOF_Obj.comp] = np.fromfile(f, dtype='S8', count=1)
    [EOF_Obj.nbodies] = np.fromfile(f, dtype='i4', count=1)
    [EOF_Obj.eof_file] = np.fromfile(f, dtype='S100', count=1)
    [EOF_Obj.mmax, EOF_Obj.nmax] = np.fromfile(f, dtype='i4', count=2)
    cosine_flat = np.fromfile(f, dtype='f8', count=(EOF_Obj.mmax + 1) * EOF_Obj.nmax)
    sine_flat = np.fromfile(f, dtype='f8', count=(EOF_Obj.mmax + 1) * EOF_Obj.nmax)
    EOF_Obj.cos = cosine_flat.reshape([(EOF_Obj.mmax + 1), EOF_Obj.nmax])
    EOF_Obj.sin = sine_flat.reshape([(EOF_Obj.mmax + 1), EOF_Obj.nmax])
    return EOF_Obj


def restore_eof_coefficients(infile):
    '''eof.py:1709-1737: (last EOF_Object, OrderedDict time -> EOF_Object)'''
    EOF_Dict = OrderedDict()
    EOF_Out = None
    with open(infile, 'rb') as f:
        [ndumps] = np.fromfile(f, dtype='i4', count=1)
        f.seek(4)
        for step in range(0, ndumps):
            try:
                EOF_Out = extract_eof_coefficients(f)
                EOF_Dict[EOF_Out.time] = EOF_Out
            except Exception:
                pass
    return EOF_Out, EOF_Dict


# ---------------------------------------------------------------------------
# coefficient time-series consumers -- eof.py:1920-2108 (host NumPy post-processing of the (time -> EOF_Object)
# dictionaries that restore_eof_coefficients / a series of accumulations produce)
# ---------------------------------------------------------------------------
def reorganize_eof_dict(EOFDict):
    '''
    eof.reorganize_eof_dict (eof.py:1920-1958): {'time', 'total'[m] (t, n) = cos^2 + sin^2, 'sum'[m] (t,),
    'cos'[m] (n, t), 'sin'[m] (n, t)}, all in time order.  The basis size is read from EOFDict[0] as in the
    reference (the dictionary must have a key 0).
    '''
    mmax, nmax = EOFDict[0].mmax, EOFDict[0].nmax
    keys = list(EOFDict.keys())
    cos = np.array([np.asarray(EOFDict[k].cos, dtype=np.float64)[:mmax + 1, :nmax] for k in keys])   # (t, m, n)
    sin = np.array([np.asarray(EOFDict[k].sin, dtype=np.float64)[:mmax + 1, :nmax] for k in keys])
    times = np.array([EOFDict[k].time for k in keys], dtype=np.float64)
    order = times.argsort()
    power = cos ** 2. + sin ** 2.
    CDict = {'time': times[order], 'total': {}, 'sum': {}, 'cos': {}, 'sin': {}}
    for mm in range(mmax + 1):
        CDict['total'][mm] = power[order, mm, :]
        CDict['sum'][mm] = np.sum(power[:, mm, :], axis=1)[order]
        CDict['cos'][mm] = cos[order, mm, :].T          # (n, t): the reference's mixed index + .T (eof.py:1955-1956)
        CDict['sin'][mm] = sin[order, mm, :].T
    return CDict


def calculate_eof_phase(EOFDict, filter=True, smooth_box=101, smooth_order=2, tol=-1.5 * np.pi, nonan=False,
                        signal_threshold=0.005):
    '''
    eof.calculate_eof_phase (eof.py:1961-2108): per-(m, n) phase arctan2(sin, cos), relative amplitude ('signal',
    normalised by sum |cos[0]|), unwrapped phase and its finite-difference rate ('speed'), the same for the
    n-summed coefficients ('netphase', 'netspeed'), and the fraction of increasing phase steps ('direction').
    Kept from the reference: the +pi in-place shift that unwrap_phase applies to a series with negative values
    also lands in DC['phase'] / DC['netphase']; |unwrapped phase| is differenced; the first speed equals the second.
    '''
    from ..utils import utils
    keys = list(EOFDict.keys())
    first = EOFDict[keys[0]]
    mmax, nmax = first.mmax, first.nmax
    cos = np.array([np.asarray(EOFDict[k].cos, dtype=np.float64)[:mmax + 1, :nmax] for k in keys])   # (t, m, n)
    sin = np.array([np.asarray(EOFDict[k].sin, dtype=np.float64)[:mmax + 1, :nmax] for k in keys])
    times = np.array([EOFDict[k].time for k in keys], dtype=np.float64)
    order = times.argsort()
    nt = len(keys)
    with np.errstate(invalid='ignore', divide='ignore'):
        phases = np.arctan2(sin, cos)
        signal = np.sqrt(cos * cos + sin * sin) / np.sum(np.sqrt(cos[:, 0, :] * cos[:, 0, :]), axis=1)[:, None, None]
    netph = np.arctan2(np.sum(sin, axis=2), np.sum(cos, axis=2))
    DC = {'time': times[order], 'phase': {}, 'netphase': {}, 'unphase': {}, 'speed': {}, 'netspeed': {},
          'signal': {}, 'direction': {}}
    dtime = np.ediff1d(DC['time'], to_begin=100.)
    for mm in range(1, mmax + 1):
        DC['phase'][mm] = phases[order, mm, :].copy()
        DC['netphase'][mm] = netph[order, mm].copy()
        DC['signal'][mm] = signal[order, mm, :].copy()
    for mm in range(1, mmax + 1):
        DC['speed'][mm] = np.zeros([nt, nmax])
        DC['unphase'][mm] = np.zeros([nt, nmax])
        DC['direction'][mm] = np.zeros(nmax)
        for nn in range(nmax):
            col = DC['phase'][mm][:, nn]                    # a view: unwrap_phase may shift it by +pi in place
            DC['direction'][mm][nn] = float(np.where(np.ediff1d(col) > 0.)[0].size) / float(col.size)
            tmp_unphase = np.abs(utils.unwrap_phase(DC['time'], col))
            if not nonan:
                good = DC['signal'][mm][:, nn] > signal_threshold
                bad = DC['signal'][mm][:, nn] < signal_threshold
                DC['unphase'][mm][good, nn] = tmp_unphase[good]
                DC['unphase'][mm][bad, nn] = np.nan
            else:
                DC['unphase'][mm][:, nn] = tmp_unphase
            series = DC['unphase'][mm][:, nn]
            if filter:
                series = utils.savitzky_golay(series, smooth_box, smooth_order)
            with np.errstate(invalid='ignore', divide='ignore'):
                DC['speed'][mm][:, nn] = np.ediff1d(series, to_begin=0.) / dtime
            DC['speed'][mm][0, nn] = DC['speed'][mm][1, nn]
        net = utils.unwrap_phase(DC['time'], DC['netphase'][mm])
        if filter:
            net = utils.savitzky_golay(net, smooth_box, smooth_order)
        with np.errstate(invalid='ignore', divide='ignore'):
            DC['netspeed'][mm] = np.ediff1d(net, to_begin=0.) / dtime
        DC['netspeed'][mm][0] = DC['netspeed'][mm][1]
    return DC
