"""
bench_configs.py -- the five configurations of BASELINE.json:configs on this rank's GPU(s).

    python profiles/bench_configs.py [--scale S]            # 1 GPU
    torchrun --nproc-per-node N profiles/bench_configs.py   # particles / orbits sharded over N GPUs

Prints one JSON line per configuration (rank 0).  Inputs are resident in HBM; CUDA events; max over ranks.
--scale S multiplies particle / orbit counts (default 1.0 = BASELINE.json's sizes divided over the ranks).
Not the headline benchmark (that is bench.py = configs[1]); this is the coverage run for the other configs.
"""
import argparse
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                    # noqa: E402
import torch.distributed as dist                # noqa: E402
from exptool_b200 import ops, parallel, synthetic as S     # noqa: E402
from oracle import oracle_np as O               # noqa: E402  (table geometry helpers only)
import bench                                    # noqa: E402


def sl_handle(lmax):
    ps, ev, ef = S.make_sl_tables(dict(lmax=lmax))
    with tempfile.TemporaryDirectory() as tmp:
        mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
        A = np.genfromtxt(mf, comments='!', skip_header=5)
    xi, r, p0, d0 = O.sl_init_table(A[:, 0], A[:, 1], A[:, 3], ps['numr'], ps['rmin'], ps['rmax'], ps['cmap'], ps['scale'])
    return ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)


def eof_handle():
    p, T, g = bench.eof_setup()
    return ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'],
                         g['numy'], g['ascale'], g['hscale'], g['cmap'], rforceC=T['rforceC'], zforceC=T['zforceC'],
                         rforceS=T['rforceS'], zforceS=T['zforceS'])


def timed(fn, reps, world):
    for _ in range(2):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dev_particles(kind, n, seed):
    gen = S.hernquist_halo if kind == 'halo' else S.exponential_disc
    out = []
    for lo in range(0, n, 2000000):              # generate on the host in pieces, keep on the device
        out.append([ops.dev(a) for a in gen(min(2000000, n - lo), seed + lo)])
    return [torch.cat([o[k] for o in out]) for k in range(4)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--configs', default='1,2,3,4,5')
    ap.add_argument('--table-fp32', type=int, default=0, help='1: option table_fp32 (float contracted tables in the per-point field kernels)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    want = set(int(c) for c in args.configs.split(','))
    ops.set_option('table_fp32', args.table_fp32)
    E = eof_handle()

    def emit(d):
        if rank == 0:
            d.update(n_gpus=world, dtype='f64', data='synthetic', tables='fp32' if args.table_fp32 else 'fp64')
            print(json.dumps(d), flush=True)

    # C1: SL lmax=4 nmax=18, accumulate + force eval, 1e5 Hernquist particles
    if 1 in want:
        n = max(int(1e5 * args.scale) // world, 1000)
        H = sl_handle(4)
        h = dev_particles('halo', n, 1001 + rank)
        coef = [None]

        def c1():
            coef[0] = parallel.sl_accumulate_sharded(H, *h, already_sharded=True)
            H.contract(coef[0])
            H.force(*h[:3])
        ms = timed(c1, 20, world)
        emit(dict(config='C1 SL lmax=4 nmax=18 accumulate + force eval', particles=n * world, ms=ms,
                  particles_per_s=n * world / ms * 1e3, pbe_per_s=n * world / ms * 1e3 * 450))
        del H, h

    # C2: EOF mmax=6 norder=18 accumulate + force eval, 1e6 disc particles (== bench.py's step)
    if 2 in want:
        n = max(int(1e6 * args.scale) // world, 1000)
        d = dev_particles('disc', n, 2002 + rank)

        def c2():
            E.prepare(*d)
            c, s = E.accumulate_prepared()
            buf = torch.stack([c, s]); parallel.allreduce_sum_(buf)
            E.contract(buf[0], buf[1])
            E.force_prepared()
        ms = timed(c2, 20, world)
        emit(dict(config='C2 EOF mmax=6 norder=18 accumulate + force eval', particles=n * world, ms=ms,
                  particles_per_s=n * world / ms * 1e3, pbe_per_s=n * world / ms * 1e3 * 234))
        del d

    # C3: combined halo (SL lmax=6) + disc (EOF mmax=6) force eval on 1e8 particles (half disc-like, half halo-like points)
    if 3 in want:
        n = max(int(1e8 * args.scale) // world, 1000)
        H6 = sl_handle(6)
        nd = n // 2
        pts_d = dev_particles('disc', nd, 3003 + rank)
        pts_h = dev_particles('halo', n - nd, 3503 + rank)
        c, s = E.accumulate(*[p[:1000000] for p in pts_d])
        ch = H6.accumulate(*[p[:1000000] for p in pts_h])
        E.contract(c * 0.025, s * 0.025); H6.contract(ch)
        x = torch.cat([pts_d[0], pts_h[0]]); y = torch.cat([pts_d[1], pts_h[1]]); z = torch.cat([pts_d[2], pts_h[2]])
        del pts_d, pts_h
        ms = timed(lambda: ops.field_force_cart(E, H6, x, y, z, rotpos=0.3), 3, world)
        emit(dict(config='C3 combined SL lmax=6 + EOF mmax=6 Cartesian force eval', particles=n * world, ms=ms,
                  particles_per_s=n * world / ms * 1e3, pbe_per_s=n * world / ms * 1e3 * (234 + 882)))
        del x, y, z
        # C4: leapfrog, 1e6 orbits x 1e4 steps in the frozen halo+disc field
        if 4 in want:
            norb = max(int(1e6 * args.scale) // world, 1000)
            nint = 10000 if args.scale >= 1.0 else max(int(10000 * args.scale), 100)
            dd = S.exponential_disc(norb, 4004 + rank)
            pos0 = np.stack(dd[:3])
            a = ops.field_force_cart(E, H6, pos0[0], pos0[1], pos0[2]).cpu().numpy()
            R = np.sqrt(pos0[0] ** 2 + pos0[1] ** 2) + 1e-12
            fr = ((a[0] + a[1]) * pos0[0] + (a[2] + a[3]) * pos0[1]) / R
            vc = np.sqrt(np.maximum(-R * fr, 1e-12))
            rng = np.random.default_rng(44 + rank)
            f = rng.uniform(0.6, 1.1, norb)
            vel0 = np.stack([-pos0[1] / R * vc * f, pos0[0] / R * vc * f, 0.1 * vc * rng.standard_normal(norb)])
            P0, V0 = ops.dev(pos0), ops.dev(vel0)
            res = [None]

            def c4():
                res[0] = ops.leapfrog(E, H6, P0, V0, nint, 3e-4, rotfreq=-5.0)
            # one untimed + one timed run (seconds long)
            c4()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); c4(); e1.record(); torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            st = res[0][0]
            bound = float((torch.sqrt(st[0] ** 2 + st[1] ** 2) < 1.0).double().mean().item())
            emit(dict(config='C4 leapfrog in frozen halo+disc BFE field', orbits=norb * world, steps=nint, ms=ms,
                      orbit_steps_per_s=norb * world * (nint - 1) / ms * 1e3, fraction_of_orbits_inside_R_lt_1=bound))
        del H6

    # C5: coefficient time series: 200 snapshots x 1e7 particles (1e6 disc + 9e6 halo), EOF + SL accumulate,
    #     ONE allreduce of the whole [200, ncoef] block at the end
    if 5 in want:
        H = sl_handle(4)
        nsnap = 200 if args.scale >= 1.0 else max(int(200 * args.scale), 4)
        ndisc = max(int(1e6 * args.scale) // world, 1000)
        nhalo = max(int(9e6 * args.scale) // world, 1000)
        base_d = dev_particles('disc', ndisc, 5005 + rank)
        base_h = dev_particles('halo', nhalo, 5505 + rank)

        def snaps():
            for k in range(nsnap):      # distinct snapshots: the base set rotated by a different angle each time
                ang = 0.01 * k
                ca, sa = float(np.cos(ang)), float(np.sin(ang))
                yield ((base_d[0] * ca - base_d[1] * sa, base_d[0] * sa + base_d[1] * ca, base_d[2], base_d[3]),
                       (base_h[0] * ca - base_h[1] * sa, base_h[0] * sa + base_h[1] * ca, base_h[2], base_h[3]))
        out = [None]

        def c5():
            out[0] = parallel.accumulate_series(E, H, snaps())
        c5()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c5(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        npart = (ndisc + nhalo) * world * nsnap
        emit(dict(config='C5 coefficient time series EOF + SL(lmax=4) accumulate, one allreduce', snapshots=nsnap,
                  particles_per_snapshot=(ndisc + nhalo) * world, ms=ms, particles_per_s=npart / ms * 1e3,
                  note='includes the on-device rotation that makes each snapshot distinct'))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
