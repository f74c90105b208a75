"""
torch_ops.py -- the hot path as PyTorch custom ops (namespace `exptool_b200`), thin shims over the C ABI of libbfe.so
(include/bfe.h).  BASELINE.json north_star: "the host side is Python calling PyTorch custom ops through a thin C-ABI
layer"; SURVEY.md section 8b: "called from the torch.library op shims".

    torch.ops.exptool_b200.eof_accumulate(handle, x, y, z, m, mmax, norder)            -> (2, mmax+1, norder)
    torch.ops.exptool_b200.eof_contract(handle, cos, sin, m1, m2, nuse, no_odd)        -> ()   (held by the handle)
    torch.ops.exptool_b200.eof_force(handle, x, y, z)                                  -> (6, n)  p0 p fr fp fz R
    torch.ops.exptool_b200.sl_accumulate(handle, x, y, z, m, nrow, nmax, no_odd)       -> (nrow, nmax)
    torch.ops.exptool_b200.sl_contract(handle, expcoef, l1, l2, nuse, no_odd)          -> ()
    torch.ops.exptool_b200.sl_force(handle, x, y, z)                                   -> (6, n)  pot0 pot1 potr pott potp r
    torch.ops.exptool_b200.field_force_cart(eof_handle, sl_handle, x, y, z, rotpos)    -> (8, n)
    torch.ops.exptool_b200.field_force_cyl(eof_handle, sl_handle, x, y, z, rotpos)     -> (8, n)
    torch.ops.exptool_b200.leapfrog(eof_handle, sl_handle, state6, nint, dt, rotfreq)  -> (6, norbit) end state

`handle` is the address of a bfe_eof / bfe_sl (ops.EOFTables.handle / ops.SLTables.handle): tensors carry the data, the
handle carries the device-resident tables.  All ops are CUDA-only (there is no CPU kernel: calling them with host
tensors raises), run on the current stream, and are what ops.EOFTables / ops.SLTables / ops.field_force_* / ops.leapfrog
call underneath.  Reference functions replaced: eof.accumulate eof.py:492, eof.accumulated_eval_particles :989,
spheresl.compute_coefficients_solitary spheresl.py:567, spheresl.all_eval_particles :1240,
Fields.return_forces_cart potential.py:445, return_forces_cyl :389, integrate.leapfrog_integrate integrate.py:53.
"""
import ctypes as C

import torch

from . import _lib

_NS = 'exptool_b200'
_DEF = torch.library.Library(_NS, 'DEF')
_IMPL = torch.library.Library(_NS, 'IMPL', 'CUDA')
_META = torch.library.Library(_NS, 'IMPL', 'Meta')


def _p(t):
    return C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f64c(*ts):
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError('exptool_b200 ops take contiguous float64 CUDA tensors')
    n = ts[0].numel()
    for t in ts[1:]:
        if t.numel() != n:
            raise ValueError('particle arrays differ in length')
    return n


def _reg(schema, impl, meta):
    name = schema.split('(')[0]
    _DEF.define(schema)
    _IMPL.impl(name, impl)
    if meta is not None:
        _META.impl(name, meta)


# ---------------------------------------------------------------- EOF
def _eof_accumulate(handle, x, y, z, m, mmax, norder):
    n = _f64c(x, y, z, m)
    out = torch.empty((2, mmax + 1, norder), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().bfe_eof_accumulate(C.c_void_p(handle), n, _p(x), _p(y), _p(z), _p(m), _p(out[0]), _p(out[1]), _st()))
    return out


def _eof_contract(handle, cosc, sinc, m1, m2, nuse, no_odd):
    _f64c(cosc, sinc)
    _lib.check(_lib.load().bfe_eof_contract(C.c_void_p(handle), _p(cosc), _p(sinc), m1, m2, nuse, int(no_odd), _st()))


def _eof_force(handle, x, y, z):
    n = _f64c(x, y, z)
    out = torch.empty((6, n), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().bfe_eof_force_contracted(C.c_void_p(handle), n, _p(x), _p(y), _p(z),
                                                    *[_p(out[i]) for i in range(6)], _st()))
    return out


# ---------------------------------------------------------------- SL
def _sl_accumulate(handle, x, y, z, m, nrow, nmax, no_odd):
    n = _f64c(x, y, z, m)
    out = torch.empty((nrow, nmax), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().bfe_sl_accumulate(C.c_void_p(handle), n, _p(x), _p(y), _p(z), _p(m), int(no_odd), _p(out), _st()))
    return out


def _sl_contract(handle, expcoef, l1, l2, nuse, no_odd):
    _f64c(expcoef)
    _lib.check(_lib.load().bfe_sl_contract(C.c_void_p(handle), _p(expcoef), l1, l2, nuse, int(no_odd), _st()))


def _sl_force(handle, x, y, z):
    n = _f64c(x, y, z)
    out = torch.empty((6, n), dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().bfe_sl_force_contracted(C.c_void_p(handle), n, _p(x), _p(y), _p(z),
                                                   *[_p(out[i]) for i in range(6)], _st()))
    return out


# ---------------------------------------------------------------- combined field, orbits
def _field(fn_name):
    def impl(eof_handle, sl_handle, x, y, z, rotpos):
        n = _f64c(x, y, z)
        out = torch.empty((8, n), dtype=torch.float64, device=x.device)
        _lib.check(getattr(_lib.load(), fn_name)(C.c_void_p(eof_handle), C.c_void_p(sl_handle), n, _p(x), _p(y), _p(z),
                                                 float(rotpos), _p(out), _st()))
        return out
    return impl


def _leapfrog(eof_handle, sl_handle, state6, nint, dt, rotfreq):
    _f64c(state6)
    if state6.dim() != 2 or state6.shape[0] != 6:
        raise ValueError('state6 must be (6, norbit): x y z vx vy vz')
    out = state6.clone()
    _lib.check(_lib.load().bfe_leapfrog(C.c_void_p(eof_handle), C.c_void_p(sl_handle), out.shape[1], int(nint), float(dt),
                                        float(rotfreq), _p(out), None, 1, 0, 1000, None, _st()))
    return out


def _empty(shape_fn):
    def meta(*a):
        ref = next(t for t in a if isinstance(t, torch.Tensor))
        return torch.empty(shape_fn(*a), dtype=torch.float64, device=ref.device)
    return meta


_reg('eof_accumulate(int handle, Tensor x, Tensor y, Tensor z, Tensor m, int mmax, int norder) -> Tensor',
     _eof_accumulate, _empty(lambda h, x, y, z, m, mmax, norder: (2, mmax + 1, norder)))
_reg('eof_contract(int handle, Tensor cosc, Tensor sinc, int m1, int m2, int nuse, bool no_odd) -> ()', _eof_contract,
     lambda *a: None)
_reg('eof_force(int handle, Tensor x, Tensor y, Tensor z) -> Tensor', _eof_force, _empty(lambda h, x, y, z: (6, x.numel())))
_reg('sl_accumulate(int handle, Tensor x, Tensor y, Tensor z, Tensor m, int nrow, int nmax, bool no_odd) -> Tensor',
     _sl_accumulate, _empty(lambda h, x, y, z, m, nrow, nmax, no_odd: (nrow, nmax)))
_reg('sl_contract(int handle, Tensor expcoef, int l1, int l2, int nuse, bool no_odd) -> ()', _sl_contract, lambda *a: None)
_reg('sl_force(int handle, Tensor x, Tensor y, Tensor z) -> Tensor', _sl_force, _empty(lambda h, x, y, z: (6, x.numel())))
_reg('field_force_cart(int eof_handle, int sl_handle, Tensor x, Tensor y, Tensor z, float rotpos) -> Tensor',
     _field('bfe_field_force_cart'), _empty(lambda e, s, x, y, z, r: (8, x.numel())))
_reg('field_force_cyl(int eof_handle, int sl_handle, Tensor x, Tensor y, Tensor z, float rotpos) -> Tensor',
     _field('bfe_field_force_cyl'), _empty(lambda e, s, x, y, z, r: (8, x.numel())))
_reg('leapfrog(int eof_handle, int sl_handle, Tensor state6, int nint, float dt, float rotfreq) -> Tensor', _leapfrog,
     _empty(lambda e, s, st, nint, dt, rf: tuple(st.shape)))

OPS = ('eof_accumulate', 'eof_contract', 'eof_force', 'sl_accumulate', 'sl_contract', 'sl_force', 'field_force_cart',
       'field_force_cyl', 'leapfrog')
