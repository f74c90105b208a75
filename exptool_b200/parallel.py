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


def allreduce_sum_(t):
    """In-place sum over ranks of a coefficient tensor (no-op on one rank)."""
    if not is_distributed():
        return t
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
