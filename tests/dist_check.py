"""
Multi-GPU parity check (NCCL): run as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_check.py
Each rank takes its block of one global particle set (parallel.shard_bounds), accumulates on its own GPU, and the
coefficient partials are summed with ONE allreduce; the result must equal the single-GPU coefficients of the whole
set (FP64 summation order only), and forces evaluated on the shards must equal the single-GPU forces.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from exptool_b200 import ops, parallel, synthetic as S
    from helpers import eof_tables, sl_tables, relerr
    meta = dict(eof_params={}, sl_params=dict(lmax=4), kind='smooth', seed=0)
    p, T, g = eof_tables(meta)
    E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                      g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                      rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
    ps, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
    x, y, z, m = S.exponential_disc(400003, 11)          # same global set on every rank
    xh, yh, zh, mh = S.hernquist_halo(100001, 12)
    c, s = parallel.eof_accumulate_sharded(E, x, y, z, m)
    ch = parallel.sl_accumulate_sharded(H, xh, yh, zh, mh)
    c1, s1 = E.accumulate(x, y, z, m)                     # whole set on this GPU
    ch1 = H.accumulate(xh, yh, zh, mh)
    e = max(relerr(c.cpu().numpy(), c1.cpu().numpy()), relerr(s.cpu().numpy(), s1.cpu().numpy()),
            relerr(ch.cpu().numpy(), ch1.cpu().numpy()))
    assert e < 1e-12, e
    # every rank must hold bit-identical reduced coefficients
    g0 = c.clone(); parallel.broadcast_(g0, src=0)
    assert torch.equal(g0, c)
    # force evaluation is embarrassingly parallel after the coefficient broadcast
    lo, hi = parallel.my_shard(len(x))
    E.contract(c, s)
    f = E.force(x[lo:hi], y[lo:hi], z[lo:hi])
    E.contract(c1, s1)
    f1 = E.force(x, y, z)[:, lo:hi]
    ef_ = max(relerr(f[i].cpu().numpy(), f1[i].cpu().numpy()) for i in range(6))
    assert ef_ < 1e-11, ef_
    # time series: 3 snapshots, ONE allreduce
    snaps = [((x[lo:hi] * (1 + 0.01 * k), y[lo:hi], z[lo:hi], m[lo:hi]), None) for k in range(3)]
    cs, ss, _ = parallel.accumulate_series(E, None, snaps)
    ck, sk = E.accumulate(x * 1.02, y, z, m)
    assert relerr(cs[2].cpu().numpy(), ck.cpu().numpy()) < 1e-12
    # the coefficient sum above went through the peer-memory kernel (bfe_peer_allreduce), not NCCL
    assert parallel._PEER['obj'] is not None and not parallel._PEER['failed'], 'peer-memory allreduce not in use'
    # the kernel against NCCL on random blocks; two instances on two streams, 300 back-to-back sums each
    # (exercises the parity double-buffering of the exchange slots and two collectives in flight)
    gen = torch.Generator(device='cuda'); gen.manual_seed(100 + rank)
    peers = [parallel.PeerAllreduce(1024) for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    blocks = [torch.randn(300, 702, dtype=torch.float64, device='cuda', generator=gen) for _ in range(2)]
    want = [b.clone() for b in blocks]
    for w in want:
        dist.all_reduce(w)
    torch.cuda.synchronize()
    for it in range(300):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                peers[k].allreduce_(blocks[k][it])
    torch.cuda.synchronize()
    for k in range(2):
        assert peers[k].first_failed_sequence() == 0
        err = float((blocks[k] - want[k]).abs().max() / want[k].abs().max())
        assert err < 1e-14, err
        ref = blocks[k].clone(); parallel.broadcast_(ref, src=0)
        assert torch.equal(ref, blocks[k]), 'peer sums differ between ranks'
    for pa in peers:
        pa.close()
    dist.barrier()
    if rank == 0:
        print('dist_check ok: world=%d coef err %.2e force err %.2e; peer-memory allreduce == NCCL to %.1e, identical on all ranks'
              % (world, e, ef_, err))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
