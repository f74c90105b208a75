"""
parallel.py -- multi-GPU plumbing: one process per GPU, torch.distributed (NCCL).

The path shards over independent units (particles, orbits) with no data-path
exchange.  The only collective is the sum of the partial
coefficient arrays, which replaces the reference's host-side
`np.sum(np.array(a_coeffs), axis=0)` over Pool workers (eof.py:1440,
spheresl.py:471): ONE allreduce of a <= 9 kB FP64 buffer per snapshot, or one
allreduce of the whole [snapshots, coefficients] block for a time series.

Works with backend "nccl" (GPU box) and "gloo" (CPU tests of the sharding logic;
gloo needs host tensors, so the reduce is staged through the host there).
"""
import numpy as np
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world():
    if is_distributed():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, world_size):
    """
    Block partition of n units over world_size ranks; rank 0 takes the remainder,
    exactly like eof.redistribute_particles (eof.py:1336-1354) /
    spheresl.redistribute_particles (spheresl.py:383-403).
    Returns a list of (lo, hi) per rank.
    """
    avg = int(np.floor(n / world_size))
    first = n - avg * (world_size - 1)
    out = [(0, first)]
    lo = first
    for _ in range(1, world_size):
        out.append((lo, lo + avg))
        lo += avg
    return out


def my_shard(n):
    rank, ws = world()
    return shard_bounds(n, ws)[rank]


class PeerAllreduce(object):
    """
    The coefficient sum as one kernel over NVLink peer memory (include/bfe.h: bfe_peer_*): every rank maps every
    other rank's exchange buffer through CUDA IPC; torch.distributed only carries the 64-byte handles once.
    One instance per stream of collectives (calls must come in the same order on every rank).

    Construction is COLLECTIVE and its outcome is agreed: every rank runs every collective of the set-up whatever
    happened locally, a success flag is reduced with MIN, and either all ranks get a working object or all ranks raise
    (after freeing what they had built) -- a rank-local failure can no longer leave some ranks on the peer kernel and
    the others on NCCL.
    """

    def __init__(self, ncoef_max=4096):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = world()
        self.ncoef_max = int(ncoef_max)
        self.h = None
        self.local = None
        self.opened = []
        err = None
        handle = C.create_string_buffer(64)
        try:
            local = C.c_void_p()
            _lib.check(self.lib.bfe_peer_buffer_create(self.ncoef_max, C.byref(local), handle))
            self.local = local
        except Exception as e:                                   # noqa: BLE001 -- reported after the agreement below
            err = e
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw if err is None else None)
        if err is None and all(hd is not None for hd in handles):
            try:
                ptrs = (C.c_void_p * self.world)()
                for r in range(self.world):
                    if r == self.rank:
                        ptrs[r] = self.local.value
                    else:
                        p = C.c_void_p()
                        _lib.check(self.lib.bfe_peer_buffer_open(handles[r], C.byref(p)))
                        self.opened.append(p)
                        ptrs[r] = p.value
                h = C.c_void_p()
                _lib.check(self.lib.bfe_peer_create(self.rank, self.world, self.ncoef_max, ptrs, C.byref(h)))
                self.h = h
            except Exception as e:                               # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError('a peer rank could not create its exchange buffer')
        ok = torch.tensor([0.0 if err is not None else 1.0], device='cuda' if dist.get_backend() == 'nccl' else 'cpu')
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)                # all ranks or none; also orders "zeroed and mapped" before the first push
        if float(ok.item()) == 0.0:
            self.close()
            raise RuntimeError('PeerAllreduce: set-up failed on at least one rank (%s)' % (err if err is not None else 'a peer failed'))

    def allreduce_(self, t, stream=None):
        """in-place sum over ranks of a contiguous FP64 device tensor with numel <= ncoef_max"""
        import ctypes as C
        from . import _lib
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= self.ncoef_max):
            raise ValueError('PeerAllreduce needs a contiguous FP64 device tensor of at most %d values' % self.ncoef_max)
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(self.lib.bfe_peer_allreduce(self.h, C.c_void_p(t.data_ptr()), t.numel(), C.c_void_p(st)))
        return t

    def first_failed_sequence(self):
        import ctypes as C
        from . import _lib
        v = C.c_uint64(0)
        _lib.check(self.lib.bfe_peer_error(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(v)))
        return int(v.value)

    def close(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        if self.h:
            self.lib.bfe_peer_destroy(self.h)
            self.h = None
        for p in self.opened:
            self.lib.bfe_peer_buffer_close(p)
        self.opened = []
        if self.local is not None:
            self.lib.bfe_peer_buffer_destroy(self.local)
            self.local = None


_PEER = {'obj': None, 'failed': False}


def _peer_allreduce_for(t):
    """the process-wide PeerAllreduce for small coefficient blocks on the current stream, or None (-> NCCL).
    Every rank takes the same branch: the eligibility tests depend only on the (rank-identical) call, and the
    constructor agrees on its outcome collectively."""
    import os
    if _PEER['failed'] or os.environ.get('BFE_PEER_ALLREDUCE', '1') == '0':
        return None
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and t.numel() <= 4096):
        return None
    if dist.get_backend() != 'nccl':
        return None
    if _PEER['obj'] is None:
        try:
            _PEER['obj'] = PeerAllreduce(4096)
        except Exception as e:                         # no peer access / IPC somewhere on this job: NCCL does the sum, on ALL ranks
            _PEER['failed'] = True
            import sys
            print('exptool_b200.parallel: peer-memory allreduce unavailable (%s); using NCCL' % (e,), file=sys.stderr)
            return None
    return _PEER['obj']


def raise_if_poisoned(values, what='coefficients'):
    """A coefficient sum that could not be formed (a rank missing for 20 s in the peer-memory kernel) comes back as NaN
    on every rank: turn that into an exception at the point where the values reach the host."""
    if is_distributed() and np.isnan(np.asarray(values)).any():
        seq = 0
        if _PEER['obj'] is not None:
            try:
                seq = _PEER['obj'].first_failed_sequence()
            except Exception:                          # noqa: BLE001
                pass
        if seq:
            raise RuntimeError('exptool_b200.parallel: the %s allreduce failed (peer-memory collective %d timed out: a rank '
                               'did not arrive within 20 s); all ranks received NaN' % (what, seq))
    return values


# ---------------------------------------------------------------------------------------------------------------------
# The reference-named entry points (eof.make_coefficients_multi, eof.compute_coefficients, spheresl.compute_coefficients)
# are RANK-LOCAL by default, like eof.accumulate and compute_coefficients_solitary: under torch.distributed every rank
# computes the coefficients of exactly the arrays it was given (one rank per snapshot of a series just works).
# Sharding the GIVEN arrays over the ranks with an allreduce of the partial sums is opt-in -- `with parallel.sharded():`
# or parallel.set_sharded(True) -- and then every rank must pass the SAME full particle set and make the same calls in
# the same order; the particle count is checked across ranks before slicing.  The explicit entry points below
# (eof_accumulate_sharded, eof_accumulate_host, sl_accumulate_sharded) always take part in the collective.
# ---------------------------------------------------------------------------------------------------------------------
_SHARDED = {'on': False}


def set_sharded(on=True):
    _SHARDED['on'] = bool(on)


def sharded_api():
    return bool(_SHARDED['on']) and is_distributed()


class sharded(object):
    """context manager: the reference-named coefficient functions shard their particle arrays over the ranks"""

    def __enter__(self):
        self.prev = _SHARDED['on']
        _SHARDED['on'] = True
        return self

    def __exit__(self, *exc):
        _SHARDED['on'] = self.prev
        return False


def _check_same_count(n):
    """all ranks must have been handed the same particle set before it is block-partitioned"""
    t = torch.tensor([float(n), -float(n)], dtype=torch.float64)
    if dist.get_backend() == 'nccl':
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if float(t[0].item()) != float(n) or -float(t[1].item()) != float(n):
        raise ValueError('exptool_b200.parallel: sharded accumulation needs the SAME particle arrays on every rank '
                         '(this rank has %d particles, the ranks hold between %d and %d); pass already_sharded=True for '
                         'per-rank shards' % (n, int(-t[1].item()), int(t[0].item())))


def allreduce_sum_(t):
    """In-place sum over ranks of a coefficient tensor (no-op on one rank).  Small FP64 device blocks go through
    the peer-memory kernel (PeerAllreduce) on NCCL jobs, everything else through torch.distributed."""
    if not is_distributed():
        return t
    peer = _peer_allreduce_for(t)
    if peer is not None:
        return peer.allreduce_(t)
    if dist.get_backend() == 'gloo' and t.is_cuda:
        h = t.cpu()
        dist.all_reduce(h, op=dist.ReduceOp.SUM)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_(t, src=0):
    """Coefficient broadcast before force evaluation / orbit integration."""
    if not is_distributed():
        return t
    if dist.get_backend() == 'gloo' and t.is_cuda:
        h = t.cpu()
        dist.broadcast(h, src=src)
        t.copy_(h)
    else:
        dist.broadcast(t, src=src)
    return t


def _slice(a, lo, hi):
    return a[lo:hi]


def eof_accumulate_sharded(E, x, y, z, m, already_sharded=False):
    """
    EOF coefficients of the GLOBAL particle set: each rank accumulates its block
    (or, if already_sharded, the arrays it was given) and the (2, mmax+1, norder)
    partials are summed with one allreduce.  Returns (cos, sin) device tensors.
    """
    if is_distributed() and not already_sharded:
        _check_same_count(len(x))
        lo, hi = my_shard(len(x))
        x, y, z, m = _slice(x, lo, hi), _slice(y, lo, hi), _slice(z, lo, hi), _slice(m, lo, hi)
    c, s = E.accumulate(x, y, z, m)
    if is_distributed():
        buf = torch.stack([c, s])
        allreduce_sum_(buf)
        c, s = buf[0], buf[1]
    return c, s


def eof_accumulate_host(E, x, y, z, m, already_sharded=False):
    """
    eof_accumulate_sharded for HOST particle arrays, returning NumPy (cos, sin): this rank's block goes through
    the chunked copy/compute pipeline (ops.EOFTables.accumulate_host) and the allreduce runs on the device
    before the single small copy out.
    """
    if is_distributed() and not already_sharded:
        _check_same_count(len(x))
        lo, hi = my_shard(len(x))
        x, y, z, m = _slice(x, lo, hi), _slice(y, lo, hi), _slice(z, lo, hi), _slice(m, lo, hi)
    return E.accumulate_host(x, y, z, m, reduce=allreduce_sum_ if is_distributed() else None)


def sl_accumulate_sharded(H, x, y, z, m, no_odd=False, already_sharded=False):
    """SL coefficients of the global particle set (see eof_accumulate_sharded)."""
    if is_distributed() and not already_sharded:
        _check_same_count(len(x))
        lo, hi = my_shard(len(x))
        x, y, z, m = _slice(x, lo, hi), _slice(y, lo, hi), _slice(z, lo, hi), _slice(m, lo, hi)
    c = H.accumulate(x, y, z, m, no_odd=no_odd)
    allreduce_sum_(c)
    return c


def sl_accumulate_host(H, x, y, z, m, no_odd=False, already_sharded=False):
    """sl_accumulate_sharded for HOST particle arrays, returning NumPy expcoef (chunked copy / compute pipeline,
    ops.SLTables.accumulate_host; the allreduce runs on the device before the one small copy out)."""
    if is_distributed() and not already_sharded:
        _check_same_count(len(x))
        lo, hi = my_shard(len(x))
        x, y, z, m = _slice(x, lo, hi), _slice(y, lo, hi), _slice(z, lo, hi), _slice(m, lo, hi)
    return H.accumulate_host(x, y, z, m, no_odd=no_odd, reduce=allreduce_sum_ if is_distributed() else None)


def accumulate_series(E, H, snapshots, no_odd=False):
    """
    Coefficient time series (BASELINE.json configs[4]): `snapshots` yields, per
    snapshot, ((xd,yd,zd,md), (xh,yh,zh,mh)) -- THIS RANK's shard of the disc and halo
    particles.  All snapshots are accumulated back to back on the stream into one
    [S, 2*(mmax+1)*norder + (lmax+1)^2*nmax] device buffer and reduced with a single
    allreduce at the end (latency-bound 9 kB reductions batched).
    Returns (cos [S,M,N], sin [S,M,N], expcoef [S,K,N]).
    """
    rows = []
    for disc, halo in snapshots:
        parts = []
        if E is not None and disc is not None:
            c, s = E.accumulate(*disc)
            parts += [c.reshape(-1), s.reshape(-1)]
        if H is not None and halo is not None:
            parts.append(H.accumulate(*halo, no_odd=no_odd).reshape(-1))
        rows.append(torch.cat(parts))
    buf = torch.stack(rows)
    allreduce_sum_(buf)
    S = buf.shape[0]
    o = 0
    cos = sin = coef = None
    if E is not None:
        k = (E.mmax + 1) * E.norder
        cos = buf[:, o:o + k].reshape(S, E.mmax + 1, E.norder); o += k
        sin = buf[:, o:o + k].reshape(S, E.mmax + 1, E.norder); o += k
    if H is not None:
        k = H.nrow * H.nmax
        coef = buf[:, o:o + k].reshape(S, H.nrow, H.nmax)
    return cos, sin, coef
