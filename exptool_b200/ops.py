"""
ops.py -- device-tensor operators over the C ABI (include/bfe.h).

PyTorch supplies device memory, the current CUDA stream and (in parallel.py)
torch.distributed; all arithmetic is in libbfe.so.  Inputs are FP64 CUDA
tensors (SoA particle arrays); NumPy arrays / CPU tensors are copied to the
current device.  Nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import torch_ops  # noqa: F401  (registers torch.ops.exptool_b200.*, which the methods below call)

# the registered custom ops, resolved once (attribute lookup + overload resolution cost ~5 us per call otherwise)
_OP_EOF_ACCUMULATE = torch.ops.exptool_b200.eof_accumulate.default
_OP_EOF_CONTRACT = torch.ops.exptool_b200.eof_contract.default
_OP_EOF_FORCE = torch.ops.exptool_b200.eof_force.default
_OP_FIELD_FORCE_CART = torch.ops.exptool_b200.field_force_cart.default
_OP_FIELD_FORCE_CYL = torch.ops.exptool_b200.field_force_cyl.default
_OP_SL_ACCUMULATE = torch.ops.exptool_b200.sl_accumulate.default
_OP_SL_CONTRACT = torch.ops.exptool_b200.sl_contract.default
_OP_SL_FORCE = torch.ops.exptool_b200.sl_force.default


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('exptool_b200 needs a CUDA device (sm_100a); there is no CPU path')


def dev(a, device=None):
    """FP64 contiguous CUDA tensor from tensor / ndarray / scalar (no copy if already so)."""
    _require_cuda()
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device())
    if isinstance(a, torch.Tensor):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))
    if t.dtype != torch.float64 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float64).contiguous()
    return t


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_host(t):
    """Device tensor -> NumPy array through a pinned staging buffer (async copy + one sync)."""
    h = torch.empty(t.shape, dtype=t.dtype, device='cpu', pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return h.numpy()


def set_option(name, value):
    """bfe_set_option: 'eof_accumulate_mode' / 'eof_force_mode' (0 auto, 1 direct, 2 sorted), 'sort_min_particles',
    'staged_eval' / 'blk_eval' (per-point kernel variants), 'table_fp32' (1: the contracted tables of the per-point
    field kernels -- Fields.return_forces_*, leapfrog, SL evaluation -- are held as float: ~1e-7 relative, faster)."""
    _lib.check(_lib.load().bfe_set_option(name.encode(), int(value)))


def get_option(name):
    """bfe_get_option: current value of a runtime option (so that a caller can restore what it changes)."""
    v = int(_lib.load().bfe_get_option(name.encode()))
    if v == -2147483648:
        raise KeyError(name)
    return v


class table_precision(object):
    """Deprecated no-op context manager (round 1 toggled the process-wide option around each call, which raced between
    threads and dropped a user's own setting).  Table precision is now a property of the HANDLE:
    EOFTables.set_table_fp32 / SLTables.set_table_fp32 (bfe_eof_set_table_fp32)."""

    def __init__(self, fp32=False):
        self.fp32 = bool(fp32)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def fp64_peak(kind='dfma'):
    """Measured FP64 peak of the current device in TFLOP/s (bfe_fp64_peak): 'dfma' (vector pipe) or 'dmma' (FP64 tensor
    cores).  ~10 ms of register-only arithmetic; synchronises the current stream."""
    _require_cuda()
    v = C.c_double(0.0)
    _lib.check(_lib.load().bfe_fp64_peak(0 if kind == 'dfma' else 1, C.byref(v), _stream()))
    return float(v.value)


def kernel_time_ms(name):
    """Duration of the latest launch of kernel `name` recorded while option 'time_kernels' was on (ms; < 0: none)."""
    return float(_lib.load().bfe_kernel_time_ms(name.encode()))


def launch_count():
    return int(_lib.load().bfe_launch_count())


class EOFTables(object):
    """
    Device-resident EOF basis (eof.parse_eof tables + eof.set_table_params geometry).

    tables: potC, rforceC, zforceC, potS, rforceS, zforceS, each
    (mmax+1, norder, numx+1, numy+1) as in eof.py:224-313.  rforce*/zforce* may be
    None if only accumulation is needed.
    """

    def __init__(self, potC, potS, mmax, norder, XMIN, dX, YMIN, dY, numx, numy, ascale, hscale, cmap,
                 rforceC=None, zforceC=None, rforceS=None, zforceS=None, dens=0):
        _require_cuda()
        self.lib = _lib.load()
        self.mmax, self.norder, self.numx, self.numy = int(mmax), int(norder), int(numx), int(numy)
        self.cmap = int(cmap)
        shape = (self.mmax + 1, self.norder, self.numx + 1, self.numy + 1)
        tabs = []
        for t in (potC, rforceC, zforceC, potS, rforceS, zforceS):
            if t is None:
                tabs.append(None)
                continue
            if tuple(t.shape) != shape:
                if t.shape[0] < shape[0] or t.shape[1] < shape[1] or tuple(t.shape[2:]) != shape[2:]:
                    raise ValueError('EOF table shape %s does not match %s' % (tuple(t.shape), shape))
                t = t[:shape[0], :shape[1]]            # MMAX/NMAX slicing as eof.force_eval, eof.py:784-793
            tabs.append(dev(t))
        self.params = _lib.EofParams(self.mmax, self.norder, self.numx, self.numy, self.cmap, int(dens),
                                     float(XMIN), float(dX), float(YMIN), float(dY), float(ascale), float(hscale))
        h = C.c_void_p()
        _lib.check(self.lib.bfe_eof_create(C.byref(self.params), *[_ptr(t) for t in tabs], _stream(), C.byref(h)))
        torch.cuda.current_stream().synchronize()   # staging tensors may now be freed
        self.h = h
        self.has_force = all(t is not None for t in tabs)
        self.device = torch.device('cuda', torch.cuda.current_device())

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.bfe_eof_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def handle(self):
        """address of the bfe_eof (the `handle` argument of the torch.ops.exptool_b200.* custom ops)"""
        return int(self.h.value)

    def set_table_fp32(self, value):
        """Table precision of the per-point field kernels for THIS handle (bfe_eof_set_table_fp32): True / False, or None to
        follow the process option 'table_fp32'.  The combined-field calls (field_force_*, leapfrog) read this handle."""
        _lib.check(self.lib.bfe_eof_set_table_fp32(self.h, -1 if value is None else int(bool(value))))

    def get_pot(self, r, z, fac=1.0):
        """eof.get_pot (eof.py:430-457): Vc, Vs as (mmax+1, norder, n) device tensors."""
        r, z = dev(r), dev(z)
        n = r.numel()
        Vc = torch.empty((self.mmax + 1, self.norder, n), dtype=torch.float64, device=self.device)
        Vs = torch.empty_like(Vc)
        _lib.check(self.lib.bfe_eof_get_pot(self.h, n, _ptr(r), _ptr(z), float(fac), _ptr(Vc), _ptr(Vs), _stream()))
        return Vc, Vs

    def clone(self):
        """
        A second handle on the same device tables with its own contraction, workspaces and counters
        (bfe_eof_clone): two particle sets can then be in flight on two streams.  Keeps the parent alive.
        """
        import copy
        other = copy.copy(self)
        other.h = None                 # never let a half-built copy destroy the parent's handle
        h = C.c_void_p()
        _lib.check(self.lib.bfe_eof_clone(self.h, _stream(), C.byref(h)))
        other.h = h
        other._parent = self
        other._prepared_n = None
        return other

    def accumulate(self, x, y, z, m):
        """eof.accumulate (eof.py:492-551) -> (cos, sin) device tensors (mmax+1, norder)."""
        x, y, z, m = dev(x), dev(y), dev(z), dev(m)
        n = x.numel()
        if not (y.numel() == n and z.numel() == n and m.numel() == n):
            raise ValueError('particle arrays differ in length')
        out = _OP_EOF_ACCUMULATE(self.handle, x, y, z, m, self.mmax, self.norder)
        return out[0], out[1]

    # -- one cell sort shared by accumulate and force evaluation on the same particles
    def prepare(self, x, y, z, m=None):
        """Counting-sort the particle set by table cell into the handle's workspace."""
        x, y, z = dev(x), dev(y), dev(z)
        m = dev(m) if m is not None else None
        n = x.numel()
        _lib.check(self.lib.bfe_eof_prepare(self.h, n, _ptr(x), _ptr(y), _ptr(z), _ptr(m), _stream()))
        self._prepared_n = n

    def accumulate_prepared(self):
        out = torch.empty((2, self.mmax + 1, self.norder), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.bfe_eof_accumulate_prepared(self.h, _ptr(out[0]), _ptr(out[1]), _stream()))
        return out[0], out[1]

    def force_prepared(self, out=None):
        n = self._prepared_n
        if out is None:
            out = torch.empty((6, n), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.bfe_eof_force_prepared(self.h, *[_ptr(out[i]) for i in range(6)], _stream()))
        return out

    def contract(self, cosc, sinc, m1=0, m2=1000, nuse=None, no_odd=False):
        cosc, sinc = self._coef(cosc), self._coef(sinc)
        nuse = self.norder if nuse is None else int(nuse)
        _OP_EOF_CONTRACT(self.handle, cosc, sinc, int(m1), int(min(m2, self.mmax)), nuse, bool(no_odd))

    def _coef(self, c):
        c = dev(c)
        if tuple(c.shape) != (self.mmax + 1, self.norder):
            if c.dim() != 2 or c.shape[0] < self.mmax + 1 or c.shape[1] < self.norder:
                raise ValueError('coefficient shape %s, expected at least %s' %
                                 (tuple(c.shape), (self.mmax + 1, self.norder)))
            c = c[:self.mmax + 1, :self.norder].contiguous()
        return c

    def force(self, x, y, z):
        """eof.accumulated_eval_particles outputs p0, p, fr, fp, fz, R (uses the held contraction)."""
        x, y, z = dev(x), dev(y), dev(z)
        return _OP_EOF_FORCE(self.handle, x, y, z)

    # -- host-array entry points (chunked copy / compute pipeline inside libbfe, bfe_host.cu)
    def accumulate_host(self, x, y, z, m, reduce=None):
        """
        eof.accumulate for HOST particle arrays -> (cos, sin) NumPy (mmax+1, norder).  `reduce`, if given, is
        applied to the (2, mmax+1, norder) device tensor before it is copied out (multi-GPU allreduce).
        Device tensors are accepted too (one-shot path).
        """
        hx, hy, hz, hm = [_host_tensor(a) for a in (x, y, z, m)]
        n = hx.numel()
        if not (hy.numel() == n and hz.numel() == n and hm.numel() == n):
            raise ValueError('particle arrays differ in length')
        buf = torch.empty((2, self.mmax + 1, self.norder), dtype=torch.float64, device=self.device)
        if hx.is_cuda:
            dx, dy, dz, dm = dev(hx), dev(hy), dev(hz), dev(hm)
            _lib.check(self.lib.bfe_eof_accumulate(self.h, n, _ptr(dx), _ptr(dy), _ptr(dz), _ptr(dm),
                                                   _ptr(buf[0]), _ptr(buf[1]), _stream()))
        else:
            _lib.check(self.lib.bfe_eof_accumulate_host(self.h, n, _ptr(hx), _ptr(hy), _ptr(hz), _ptr(hm),
                                                        _ptr(buf[0]), _ptr(buf[1]), _stream()))
        if reduce is not None:
            reduce(buf)
        h = to_host(buf)               # synchronises the stream: the host inputs may be released after this
        return h[0], h[1]

    def force_host(self, x, y, z):
        """
        eof.accumulated_eval_particles for HOST particle arrays -> (6, n) NumPy array p0, p, fr, fp, fz, R
        (uses the held contraction).  The result lives in pinned memory from torch's caching host allocator.
        """
        hx, hy, hz = [_host_tensor(a) for a in (x, y, z)]
        n = hx.numel()
        if hx.is_cuda:
            return to_host(self.force(hx, hy, hz))
        h = pinned_empty((6, n))
        _lib.check(self.lib.bfe_eof_force_host(self.h, n, _ptr(hx), _ptr(hy), _ptr(hz),
                                               *[_ptr(h[i]) for i in range(6)], _stream()))
        torch.cuda.current_stream().synchronize()
        return h.numpy()

    def force_eval_points(self, r, z, phi):
        """eof.force_eval outputs fr, fp, fz, p (incl. m=0), p0 at cylindrical points."""
        r, z, phi = dev(r), dev(z), dev(phi)
        n = r.numel()
        out = torch.empty((5, n), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.bfe_eof_force_eval_points(self.h, n, _ptr(r), _ptr(z), _ptr(phi),
                                                      *[_ptr(out[i]) for i in range(5)], _stream()))
        return out


# ---------------------------------------------------------------------------
# host-array helpers (the reference-facing API takes and returns HOST arrays; the chunked
# H2D | kernels | D2H pipeline itself lives in libbfe: bfe_eof_accumulate_host / bfe_eof_force_host)
# ---------------------------------------------------------------------------
def _host_tensor(a):
    """FP64 contiguous tensor view of ndarray / tensor (copy only if dtype/layout requires)."""
    if isinstance(a, torch.Tensor):
        t = a
        if t.dtype != torch.float64 or not t.is_contiguous():
            t = t.to(torch.float64).contiguous()
        return t
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)))


def pinned_empty(shape, dtype=torch.float64):
    """Pinned host tensor from torch's caching host allocator (first use pays cudaHostAlloc, later ones are free)."""
    return torch.empty(shape, dtype=dtype, device='cpu', pin_memory=True)


class SLTables(object):
    """Device-resident SL basis (halo_methods.read_cached_table + init_table)."""

    def __init__(self, lmax, nmax, numr, cmap, scale, evtable, eftable, xi, p0, d0=None):
        _require_cuda()
        self.lib = _lib.load()
        self.lmax, self.nmax, self.numr, self.cmap = int(lmax), int(nmax), int(numr), int(cmap)
        self.scale = float(scale)
        ev = np.asarray(evtable)[:self.lmax + 1, :self.nmax] if not isinstance(evtable, torch.Tensor) \
            else evtable[:self.lmax + 1, :self.nmax]
        ef = np.asarray(eftable)[:self.lmax + 1, :self.nmax] if not isinstance(eftable, torch.Tensor) \
            else eftable[:self.lmax + 1, :self.nmax]
        ev, ef, xi, p0 = dev(ev), dev(ef), dev(xi), dev(p0)
        d0 = dev(d0) if d0 is not None else None
        if tuple(ef.shape) != (self.lmax + 1, self.nmax, self.numr) or xi.numel() != self.numr or p0.numel() != self.numr:
            raise ValueError('SL table shapes do not match (lmax, nmax, numr)')
        self.params = _lib.SlParams(self.lmax, self.nmax, self.numr, self.cmap, self.scale)
        h = C.c_void_p()
        _lib.check(self.lib.bfe_sl_create(C.byref(self.params), _ptr(ev), _ptr(ef), _ptr(xi), _ptr(p0), _ptr(d0),
                                          _stream(), C.byref(h)))
        torch.cuda.current_stream().synchronize()
        self.h = h
        self.nrow = (self.lmax + 1) ** 2
        self.device = torch.device('cuda', torch.cuda.current_device())

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.lib.bfe_sl_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def handle(self):
        """address of the bfe_sl (the `handle` argument of the torch.ops.exptool_b200.* custom ops)"""
        return int(self.h.value)

    def set_table_fp32(self, value):
        """Table precision of the SL-only evaluation kernels for THIS handle (bfe_sl_set_table_fp32)."""
        _lib.check(self.lib.bfe_sl_set_table_fp32(self.h, -1 if value is None else int(bool(value))))

    def accumulate(self, x, y, z, m, no_odd=False):
        """spheresl.compute_coefficients_solitary (spheresl.py:567-656) -> expcoef ((lmax+1)^2, nmax)."""
        x, y, z, m = dev(x), dev(y), dev(z), dev(m)
        n = x.numel()
        if not (y.numel() == n and z.numel() == n and m.numel() == n):
            raise ValueError('particle arrays differ in length')
        return _OP_SL_ACCUMULATE(self.handle, x, y, z, m, self.nrow, self.nmax, bool(no_odd))

    # -- host-array entry points (chunked copy / compute pipeline inside libbfe, bfe_host.cu)
    def accumulate_host(self, x, y, z, m, no_odd=False, reduce=None):
        """spheresl.compute_coefficients for HOST particle arrays -> expcoef NumPy ((lmax+1)^2, nmax).  `reduce`, if given,
        is applied to the device tensor before it is copied out (multi-GPU allreduce).  Device tensors take the one-shot path."""
        hx, hy, hz, hm = [_host_tensor(a) for a in (x, y, z, m)]
        n = hx.numel()
        if not (hy.numel() == n and hz.numel() == n and hm.numel() == n):
            raise ValueError('particle arrays differ in length')
        if hx.is_cuda:
            buf = self.accumulate(hx, hy, hz, hm, no_odd=no_odd)
        else:
            buf = torch.empty((self.nrow, self.nmax), dtype=torch.float64, device=self.device)
            _lib.check(self.lib.bfe_sl_accumulate_host(self.h, n, _ptr(hx), _ptr(hy), _ptr(hz), _ptr(hm), int(bool(no_odd)),
                                                       _ptr(buf), _stream()))
        if reduce is not None:
            reduce(buf)
        return to_host(buf)            # synchronises the stream: the host inputs may be released after this

    def force_host(self, x, y, z):
        """spheresl.all_eval_particles potential outputs for HOST arrays -> (6, n) NumPy pot0, pot1, potr, pott, potp, rr
        (uses the held contraction; pinned result buffer from torch's caching host allocator)."""
        hx, hy, hz = [_host_tensor(a) for a in (x, y, z)]
        n = hx.numel()
        if hx.is_cuda:
            return to_host(self.force(hx, hy, hz))
        h = pinned_empty((6, n))
        _lib.check(self.lib.bfe_sl_force_host(self.h, n, _ptr(hx), _ptr(hy), _ptr(hz), *[_ptr(h[i]) for i in range(6)], _stream()))
        torch.cuda.current_stream().synchronize()
        return h.numpy()

    def _coef(self, c):
        c = dev(c)
        if tuple(c.shape) != (self.nrow, self.nmax):
            if c.dim() != 2 or c.shape[0] < self.nrow or c.shape[1] < self.nmax:
                raise ValueError('expcoef shape %s, expected at least %s' % (tuple(c.shape), (self.nrow, self.nmax)))
            c = c[:self.nrow, :self.nmax].contiguous()
        return c

    def contract(self, expcoef, l1=-1000, l2=1000, nuse=None, no_odd=False):
        c = self._coef(expcoef)
        nuse = self.nmax if nuse is None else int(nuse)
        l1 = max(int(l1), 0)
        l2 = min(int(l2), self.lmax)
        _OP_SL_CONTRACT(self.handle, c, l1, l2, nuse, bool(no_odd))

    def force(self, x, y, z):
        """spheresl.all_eval_particles outputs pot0, pot1, potr, pott, potp, rr."""
        x, y, z = dev(x), dev(y), dev(z)
        return _OP_SL_FORCE(self.handle, x, y, z)

    def radial_matrices(self, r, dens=True, force=True, pot=True):
        """spheresl.get_halo_dens_pot_force (spheresl.py:106-160) at n radii: (dens, force, pot), each
        (lmax+1, nmax, n) or None when not requested."""
        r = dev(r)
        n = r.numel()
        outs = [torch.empty((self.lmax + 1, self.nmax, n), dtype=torch.float64, device=self.device) if want else None
                for want in (dens, force, pot)]
        _lib.check(self.lib.bfe_sl_radial_matrices(self.h, n, _ptr(r), *[_ptr(o) for o in outs], _stream()))
        return tuple(outs)

    def contract_density(self, expcoef, l1=-1000, l2=1000, nuse=None, no_odd=False):
        """Density rows sum_n c ef sqrt(ev) (spheresl.py:148); needs d0 at construction."""
        c = self._coef(expcoef)
        nuse = self.nmax if nuse is None else int(nuse)
        _lib.check(self.lib.bfe_sl_contract_density(self.h, _ptr(c), max(int(l1), 0), min(int(l2), self.lmax), nuse,
                                                    int(bool(no_odd)), _stream()))

    def density(self, x, y, z):
        """den0, den1 of spheresl.all_eval_particles (spheresl.py:1271,1323,1351) -> (2, n)."""
        x, y, z = dev(x), dev(y), dev(z)
        n = x.numel()
        out = torch.empty((2, n), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.bfe_sl_density_contracted(self.h, n, _ptr(x), _ptr(y), _ptr(z), _ptr(out[0]), _ptr(out[1]),
                                                      _stream()))
        return out

    def density_eval_points(self, r, costh, phi):
        """den0, den1 of spheresl.all_eval (spheresl.py:1046,1071,1092) -> (2, n)."""
        r, costh, phi = dev(r), dev(costh), dev(phi)
        n = r.numel()
        out = torch.empty((2, n), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.bfe_sl_density_eval_points(self.h, n, _ptr(r), _ptr(costh), _ptr(phi), _ptr(out[0]),
                                                       _ptr(out[1]), _stream()))
        return out

    def force_eval_points(self, r, costh, phi, trig_index_l=True):
        """spheresl.force_eval (trig_index_l) / all_eval outputs potr, pott, potp, pot1, pot0."""
        r, costh, phi = dev(r), dev(costh), dev(phi)
        n = r.numel()
        out = torch.empty((5, n), dtype=torch.float64, device=self.device)
        _lib.check(self.lib.bfe_sl_force_eval_points(self.h, n, _ptr(r), _ptr(costh), _ptr(phi),
                                                     int(bool(trig_index_l)), *[_ptr(out[i]) for i in range(5)],
                                                     _stream()))
        return out


def field_force_cart(eof_tables, sl_tables, x, y, z, rotpos=0.0):
    """Fields.return_forces_cart (potential.py:445-497) at n points -> (8, n) device tensor."""
    x, y, z = dev(x), dev(y), dev(z)
    return _OP_FIELD_FORCE_CART(eof_tables.handle, sl_tables.handle, x, y, z, float(rotpos))


def field_force_cyl(eof_tables, sl_tables, x, y, z, rotpos=0.0):
    """Fields.return_forces_cyl (potential.py:389-440) at n points -> (8, n) device tensor."""
    x, y, z = dev(x), dev(y), dev(z)
    return _OP_FIELD_FORCE_CYL(eof_tables.handle, sl_tables.handle, x, y, z, float(rotpos))


def leapfrog(eof_tables, sl_tables, pos0, vel0, nint, dt, rotfreq=0.0, traj_stride=0, apse=False, ap_max=1000):
    """
    integrate.leapfrog_integrate (integrate.py:53-190) for a batch of orbits.
    pos0, vel0: (3, norbit).  Returns (state6 (6, norbit) at the last step,
    traj (nsave, 10, norbit) or None, nsteps (norbit,) int32).
    """
    pos0, vel0 = dev(pos0), dev(vel0)
    pos0 = pos0.reshape(3, -1)
    vel0 = vel0.reshape(3, -1)
    norb = pos0.shape[1]
    state = torch.cat([pos0, vel0], dim=0).contiguous()
    traj = None
    if traj_stride and traj_stride > 0:
        nsave = (int(nint) - 1) // int(traj_stride) + 1
        traj = torch.zeros((nsave, 10, norb), dtype=torch.float64, device=state.device)
    nsteps = torch.empty((norb,), dtype=torch.int32, device=state.device)
    stride = int(traj_stride) if traj is not None else 1
    if np.ndim(dt) == 0 and not isinstance(dt, torch.Tensor):
        _lib.check(eof_tables.lib.bfe_leapfrog(eof_tables.h, sl_tables.h, norb, int(nint), float(dt), float(rotfreq),
                                               _ptr(state), _ptr(traj), stride, int(bool(apse)), int(ap_max),
                                               _ptr(nsteps), _stream()))
    else:
        dts = dev(dt).reshape(-1)                  # one step size per orbit
        if dts.numel() != norb:
            raise ValueError('dt must be a scalar or have one entry per orbit')
        _lib.check(eof_tables.lib.bfe_leapfrog_dt(eof_tables.h, sl_tables.h, norb, int(nint), _ptr(dts),
                                                  float(rotfreq), _ptr(state), _ptr(traj), stride,
                                                  int(bool(apse)), int(ap_max), _ptr(nsteps), _stream()))
    return state, traj, nsteps


# ---------------------------------------------------------------------------
# pre-accumulation transforms of a device-resident snapshot
# ---------------------------------------------------------------------------
def bar_fourier_angle(x, y, minr=0.0, maxr=1.0):
    """pattern.BarTransform.bar_fourier_compute (pattern.py:155-169): atan2(sum sin 2phi, sum cos 2phi) / 2
    over minr < R < maxr.  x, y device (or host) arrays; returns a Python float (one 16-byte copy out)."""
    x, y = dev(x), dev(y)
    out = torch.empty(2, dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().bfe_bar_fourier(x.numel(), _ptr(x), _ptr(y), float(minr), float(maxr), _ptr(out), _stream()))
    a, b = out.cpu().tolist()
    return float(np.arctan2(b, a) / 2.0)


def affine_xy(x, y, z=None, angle=0.0, center=(0.0, 0.0, 0.0), out=None):
    """(x, y) rotated counter-clockwise by `angle`, then shifted by -center; z shifted (pattern.py:118-139,
    potential.py:213-219).  Returns new device tensors (or writes `out` = (xo, yo[, zo]))."""
    x, y = dev(x), dev(y)
    z = dev(z) if z is not None else None
    if out is None:
        out = (torch.empty_like(x), torch.empty_like(y)) + ((torch.empty_like(z),) if z is not None else ())
    _lib.check(_lib.load().bfe_affine_xy(x.numel(), float(angle), float(center[0]), float(center[1]), float(center[2]),
                                         _ptr(x), _ptr(y), _ptr(z), _ptr(out[0]), _ptr(out[1]),
                                         _ptr(out[2]) if z is not None else C.c_void_p(0), _stream()))
    return out


def inner_center_of_mass(x, y, z, m, ncenter=10000, values=None):
    """Mass-weighted centre of the `ncenter` innermost particles (potential.py:158-176) -> (xc, yc, zc) floats.
    `values` = (vx, vy, vz, vm): sum these arrays at the selected indices instead (the reference ranks the disc
    and indexes the halo arrays with the result, potential.py:190-200); they must have at least len(x) entries."""
    x, y, z = dev(x), dev(y), dev(z)
    if values is None:
        vx, vy, vz, vm = x, y, z, dev(m)
    else:
        vx, vy, vz, vm = [dev(a) for a in values]
        if min(vx.numel(), vy.numel(), vz.numel(), vm.numel()) < x.numel():
            raise IndexError('inner_center_of_mass: value arrays are shorter than the ranked set')
    out = torch.empty(4, dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().bfe_inner_com(x.numel(), _ptr(x), _ptr(y), _ptr(z), _ptr(vx), _ptr(vy), _ptr(vz), _ptr(vm),
                                         int(ncenter), _ptr(out), _stream()))
    sx, sy, sz, sm = out.cpu().tolist()
    return sx / sm, sy / sm, sz / sm


def eof_return_bins(r, z, rmin, dR, zmin, dZ, numx, numy, ascale, hscale, cmap):
    """eof.return_bins (eof.py:354-427) -> X, Y (float64), ix, iy (int64) device tensors."""
    r, z = dev(r), dev(z)
    n = r.numel()
    params = _lib.EofParams(0, 1, int(numx), int(numy), int(cmap), 0, float(rmin), float(dR), float(zmin), float(dZ),
                            float(ascale), float(hscale))
    X = torch.empty(n, dtype=torch.float64, device=r.device)
    Y = torch.empty_like(X)
    ix = torch.empty(n, dtype=torch.int64, device=r.device)
    iy = torch.empty_like(ix)
    _lib.check(_lib.load().bfe_eof_return_bins(C.byref(params), n, _ptr(r), _ptr(z), _ptr(X), _ptr(Y), _ptr(ix), _ptr(iy),
                                               _stream()))
    return X, Y, ix, iy


def legendre_tables(lmax, x, derivative=True):
    """spheresl.legendre_R / dlegendre_R (spheresl.py:664-770) at n arguments: P, dP as (lmax+1, lmax+1, n)."""
    x = dev(x)
    n = x.numel()
    P = torch.empty((lmax + 1, lmax + 1, n), dtype=torch.float64, device=x.device)
    dP = torch.empty_like(P) if derivative else None
    _lib.check(_lib.load().bfe_legendre_tables(int(lmax), n, _ptr(x), _ptr(P), _ptr(dP), _stream()))
    return P, dP
