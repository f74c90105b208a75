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
    """

    def __init__(self, ncoef_max=4096):
        import ctypes as C
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world = world()
        self.ncoef_max = int(ncoef_max)
        self.h = None
        local = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(self.lib.bfe_peer_buffer_create(self.ncoef_max, C.byref(local), handle))
        self.local = local
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw)
        self.opened = []
        ptrs = (C.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = local.value
            else:
                p = C.c_void_p()
                _lib.check(self.lib.bfe_peer_buffer_open(handles[r], C.byref(p)))
                self.opened.append(p)
                ptrs[r] = p.value
        h = C.c_void_p()
        _lib.check(self.lib.bfe_peer_create(self.rank, self.world, self.ncoef_max, ptrs, C.byref(h)))
        self.h = h
        dist.barrier()                      # every buffer is zeroed and mapped before the first push

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
        if self.h:
            torch.cuda.synchronize()
            self.lib.bfe_peer_destroy(self.h)
            self.h = None
            for p in self.opened:
                self.lib.bfe_peer_buffer_close(p)
            self.lib.bfe_peer_buffer_destroy(self.local)


_PEER = {'obj': None, 'failed': False}


def _peer_allreduce_for(t):
    """the process-wide PeerAllreduce for small coefficient blocks on the current stream, or None (-> NCCL)"""
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
        except Exception as e:                         # no peer access / IPC on this box: NCCL does the sum
            _PEER['failed'] = True
            import sys
            print('exptool_b200.parallel: peer-memory allreduce unavailable (%s); using NCCL' % (e,), file=sys.stderr)
            return None
    return _PEER['obj']


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
        lo, hi = my_shard(len(x))
        x, y, z, m = _slice(x, lo, hi), _slice(y, lo, hi), _slice(z, lo, hi), _slice(m, lo, hi)
    return E.accumulate_host(x, y, z, m, reduce=allreduce_sum_ if is_distributed() else None)


def sl_accumulate_sharded(H, x, y, z, m, no_odd=False, already_sharded=False):
    """SL coefficients of the global particle set (see eof_accumulate_sharded)."""
    if is_distributed() and not already_sharded:
        lo, hi = my_shard(len(x))
        x, y, z, m = _slice(x, lo, hi), _slice(y, lo, hi), _slice(z, lo, hi), _slice(m, lo, hi)
    c = H.accumulate(x, y, z, m, no_odd=no_odd)
    allreduce_sum_(c)
    return c


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
