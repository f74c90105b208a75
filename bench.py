#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

metric   : particles/sec for BFE accumulate+force eval   (BASELINE.json:metric)
workload : configs[1] -- EOF cylindrical disc mmax=6 norder=18, coefficient accumulation +
           force evaluation on a 10^6-particle synthetic exponential disc per GPU.
step     : one pass of the hot path over one particle set:
             eof_accumulate (10^6 particles -> 2x7x18 coefficients)
             [N>1: one NCCL allreduce of the 252 coefficients]
             eof_contract   (coefficients x tables -> contracted grids)
             eof_force      (p0,p,fr,fp,fz,R at the same 10^6 particles)
value    : whole-job particles/s, inputs resident in HBM, CUDA-event time, max over ranks.
e2e      : same metric through the reference-facing Python API (eof.make_coefficients_multi +
           eof.accumulated_eval_particles) with HOST buffers: pinned host -> device copy of the
           particle arrays and device -> host copy of the six output arrays inside the timed region.
roofline : dominant kernel, algorithmic HBM bytes / CUDA-event duration vs MEASURED_PEAKS.json.
cpu_baseline : the oracle's C restatement of the reference's arithmetic (oracle/bfe_oracle.c, OpenMP,
           `kind: port`) on all host threads over a bounded sample of the same workload.

configs  : the other four BASELINE.json configurations (C1 SL 10^5, C3 combined field 10^8 points, C4 10^6 orbits x 10^4
           steps, C5 coefficient time series 200 x 10^7), each with its own `roofline` (HBM fraction of the measured copy
           bandwidth AND FP64 fraction of the measured DFMA peak; `bound` = the larger), `cpu_baseline` (a port of the
           reference formulation on a stated subsample) and `parity` (max relative error of the GPU result against that
           CPU result on the same subsample).  Sizes are BASELINE.json's totals divided over the ranks (strong scaling).
           --configs none skips them, --configs-scale S shrinks them (smoke runs).

--impl reference times that CPU port as the reference arm (the reference itself is pure
Python under /root/reference, which does not exist on the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'particles/sec for BFE accumulate+force eval'
UNIT = 'particles/s'
N_PART = 1000000
WORKLOAD = 'EOF mmax=6 norder=18 numx=128 numy=64: accumulate + force eval, 10^6-particle exponential disc per GPU'
BYTES_ACC = 32       # x,y,z,m read                          
BYTES_FORCE = 72     # x,y,z read + p0,p,fr,fp,fz,R written  
PBE_PER_PARTICLE = 234


# ----------------------------------------------------------------------------- helpers
def eof_setup():
    from exptool_b200 import synthetic as S
    from exptool_b200.basis import eof as beof
    p, T = S.make_eof_tables({}, kind='smooth')
    XMIN, XMAX, dX, YMIN, YMAX, dY = beof.set_table_params(RMAX=p['rmax'], RMIN=p['rmin'], ASCALE=p['ascale'],
                                                           HSCALE=p['hscale'], NUMX=p['numx'], NUMY=p['numy'],
                                                           CMAP=p['cmap'])
    g = dict(XMIN=float(XMIN), dX=float(dX), YMIN=float(YMIN), dY=float(dY), numx=p['numx'], numy=p['numy'], mmax=p['mmax'],
             norder=p['norder'], ascale=p['ascale'], hscale=p['hscale'], cmap=p['cmap'])
    return p, T, g


def sl_setup(lmax):
    """SL cache tables + model on the cache's xi grid, through the product's own host readers (halo_methods.init_table)"""
    import tempfile
    from exptool_b200 import synthetic as S
    from exptool_b200.utils import halo_methods
    ps, ev, ef = S.make_sl_tables(dict(lmax=lmax))
    with tempfile.TemporaryDirectory() as tmp:
        mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
        xi, r, p0, d0 = halo_methods.init_table(mf, ps['numr'], ps['rmin'], ps['rmax'], cmap=ps['cmap'], scale=ps['scale'])
    return ps, ev, ef, xi, p0, d0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.samples.append(line.strip())
        except Exception:
            pass

    def stop(self):
        self.stop_flag.set()
        try:
            if self.proc:
                self.proc.terminate()
        except Exception:
            pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            f = [t.strip() for t in s.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith('active'):
                    reasons.add(nme)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ----------------------------------------------------------------------------- CPU port (oracle)
# The reference is pure Python under /root/reference and cannot travel to the GPU box, so the CPU legs time the
# oracle's C restatement (oracle/bfe_oracle.c: the reference's direct formulation, FP64, OpenMP over particles,
# all host threads) -- a stronger baseline than the reference's own NumPy path (the NumPy port of the same
# arithmetic is timed once beside it for context).
def cpu_port_run(n, seed=7000):
    """accumulate + force eval of n particles with the C port on all host threads; wall seconds."""
    from exptool_b200 import synthetic as S
    from oracle import oracle_c as OC
    g, T = _CPU['g'], _CPU['T']
    x, y, z, m = S.exponential_disc(n, seed)
    t0 = time.perf_counter()
    c, s = OC.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g)
    OC.eof_force(x, y, z, c, s, T, g)
    return time.perf_counter() - t0


def numpy_port_rate(n=20000):
    """particles/s of the NumPy port (oracle_np) on ONE core, for context."""
    from exptool_b200 import synthetic as S
    from oracle import oracle_np as O
    g, T = _CPU['g'], _CPU['T']
    x, y, z, m = S.exponential_disc(n, 7001)
    geo = (g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'])
    t0 = time.perf_counter()
    c, s = O.eof_accumulate(x, y, z, m, T['potC'], T['potS'], g['mmax'], g['norder'], *geo, g['ascale'], g['hscale'], g['cmap'])
    O.eof_force_particles(x, y, z, c, s, T['potC'], T['rforceC'], T['zforceC'], T['potS'], T['rforceS'], T['zforceS'],
                          *geo, g['mmax'], g['norder'], g['ascale'], g['hscale'], g['cmap'])
    return n / (time.perf_counter() - t0)


_CPU = {}


def cpu_setup():
    from oracle import oracle_c as OC
    p, T, g = eof_setup()
    _CPU['g'] = g; _CPU['T'] = T
    return OC.use_all_cores()          # not OMP_NUM_THREADS: torchrun sets that to 1 for its workers


def cpu_baseline(target_seconds=12.0):
    cores = cpu_setup()
    dt = cpu_port_run(20000)                                   # calibration + page-in of the tables
    rate = 20000 / dt
    n_sample = int(min(max(rate * target_seconds, 50000), 4 * N_PART))
    dt = cpu_port_run(n_sample, seed=7002)
    return {'value': n_sample / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': '%d particles (accumulate + force eval, reference direct formulation), C port with OpenMP on '
                      '%d threads, %.1f s' % (n_sample, cores, dt),
            'numpy_port_one_core': numpy_port_rate()}


def run_reference(args):
    """Reference arm: the C port on all host threads, each step a bounded sample."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = cpu_setup()
    dt = cpu_port_run(20000)
    rate = 20000 / dt
    budget = 150.0 / max(args.steps + args.warmup, 1)             # whole run within a few minutes
    n_step = int(min(max(rate * min(budget, 10.0), 20000), N_PART))
    for k in range(args.warmup):
        cpu_port_run(n_step, seed=7100 + k)
    t = 0.0
    for k in range(args.steps):
        t += cpu_port_run(n_step, seed=7200 + k)
    done = n_step * args.steps
    value = done / t
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'particles_per_gpu': N_PART, 'sample_particles_per_step': n_step},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d particles per step (accumulate + force eval), C port of the reference '
                                       'formulation with OpenMP on %d threads' % (n_step, cores)},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- the other BASELINE configurations
FIELD_FLOP_PER_POINT = 2 * 741 + 409 + 136   # executed FP64 flop of one combined-field evaluation at mmax=6 / lmax=6: DFMA x2 + DMUL + DADD,
                                             # static SASS count of field_rec_kernel<6,6> (fully unrolled; profiles/r02_sass_field_rec_kernel.txt;
                                             # 2389 before the polynomial SL blocks: fewer flop for the same result lowers this fraction)
FIELD_BYTES_PER_POINT = 24 + 64              # x,y,z in + the 8-tuple out


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = float(np.max(np.abs(b))) if b.size else 0.0
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b))) if b.size else 0.0


def _roof(bytes_alg, flop_exec, ms, peaks, kernel, note=None):
    """roofline object of one configuration: HBM side against the measured copy bandwidth, FP64 side against the measured
    DFMA peak (bfe_fp64_peak); `bound` is whichever fraction is larger (SURVEY.md section 8d rule v)."""
    sec = ms * 1e-3
    hbm = bytes_alg / sec / 1e9 if bytes_alg else 0.0
    hbm_frac = hbm / peaks['hbm_gbs']
    out = {'kernel': kernel, 'hbm': {'achieved': hbm, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': hbm_frac,
                                     'algorithmic_bytes': bytes_alg, 'peak_source': peaks['hbm_source']}, 'traffic': None}
    if flop_exec:
        tf = flop_exec / sec / 1e12
        out['fp64'] = {'achieved': tf, 'peak': peaks['dfma_tflops'], 'unit': 'TFLOP/s', 'frac': tf / peaks['dfma_tflops'],
                       'executed_flop': flop_exec, 'peak_source': 'measured in this run: bfe_fp64_peak(DFMA); DMMA %.1f'
                       % peaks['dmma_tflops']}
    else:
        out['fp64'] = None
    if out['fp64'] and out['fp64']['frac'] > hbm_frac:
        out.update(bound='fp64', achieved=out['fp64']['achieved'], peak=peaks['dfma_tflops'], unit='TFLOP/s', frac=out['fp64']['frac'])
    else:
        out.update(bound='hbm', achieved=hbm, peak=peaks['hbm_gbs'], unit='GB/s', frac=hbm_frac)
    if note:
        out['note'] = note
    return out


def run_configs(args, E, T, g, world, rank, dev, peaks):
    """C1, C3, C4, C5 of BASELINE.json:configs on this job's GPUs (C2 is the headline step above)."""
    import torch
    import torch.distributed as dist
    from exptool_b200 import ops, parallel, synthetic as S
    want = set(c.strip().upper() for c in args.configs.split(',')) if args.configs != 'all' else {'C1', 'C3', 'C4', 'C5'}
    scale = args.configs_scale
    cpu_ok = (rank == 0 and world == 1 and not args.no_cpu)
    res = {}

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def dev_particles(kind, n, seed):
        gen = S.hernquist_halo if kind == 'halo' else S.exponential_disc
        parts = [[ops.dev(a) for a in gen(min(2000000, n - lo), seed + lo)] for lo in range(0, n, 2000000)]
        return [torch.cat([q[k] for q in parts]) for k in range(4)]

    def sl_handle(lmax):
        ps, ev, ef, xi, p0, d0 = sl_setup(lmax)
        return ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0), (ps, ev, ef, xi, p0, d0)

    def frozen(cosc, sinc, coef, sl):
        from oracle import oracle_np as O
        ps, ev, ef, xi, p0, d0 = sl
        return O.FrozenField(cos=cosc, sin=sinc, potC=T['potC'], rforceC=T['rforceC'], zforceC=T['zforceC'], potS=T['potS'],
                             rforceS=T['rforceS'], zforceS=T['zforceS'], XMIN=g['XMIN'], dX=g['dX'], YMIN=g['YMIN'], dY=g['dY'],
                             numx=g['numx'], numy=g['numy'], mmax=g['mmax'], norder=g['norder'], ascale=g['ascale'],
                             hscale=g['hscale'], cmapdisk=g['cmap'], halofac=1.0, expcoef=coef, xihalo=xi, p0halo=p0, d0halo=d0,
                             cmaphalo=ps['cmap'], scalehalo=ps['scale'], lmaxhalo=ps['lmax'], nmaxhalo=ps['nmax'],
                             evtablehalo=ev, eftablehalo=ef)

    # ---------------- C1: SL lmax=4 nmax=18, accumulate + force eval, 10^5 Hernquist particles
    if 'C1' in want:
        n = max(int(1e5 * scale) // world, 1000)
        H, sl = sl_handle(4)
        h = dev_particles('halo', n, 1001 + rank)
        box = {}

        def c1():
            box['coef'] = parallel.sl_accumulate_sharded(H, *h, already_sharded=True)
            H.contract(box['coef'])
            box['out'] = H.force(*h[:3])
        ms = timed(c1, 20)
        entry = {'config': 'C1 SL lmax=4 nmax=18: accumulate + force eval, 10^5-particle Hernquist halo', 'particles': n * world,
                 'ms': ms, 'value': n * world / ms * 1e3, 'unit': 'particles/s', 'pbe_per_s': n * world / ms * 1e3 * 450,
                 'scaling': 'strong',
                 'roofline': _roof((32 + 72) * n, None, ms, peaks, 'sl_accumulate + sl_contract + sl_force',
                                   note='10^5 particles are ~8 launches of 5-30 us: launch/latency-bound, not a bandwidth test')}
        # e2e through the reference-facing API with host buffers: spheresl.compute_coefficients (chunked host pipeline,
        # bfe_sl_accumulate_host) + spheresl.eval_particles (one upload, potential + density rows, one copy out)
        import tempfile
        from exptool_b200.basis import spheresl as bsl
        with tempfile.TemporaryDirectory() as tmpd:
            ps_, ev_, ef_ = sl[0], sl[1], sl[2]
            sf = S.write_sl_cache(os.path.join(tmpd, 'sl.cache'), ps_, ev_, ef_)
            mf = S.write_hernquist_model(os.path.join(tmpd, 'sl.model'), a=ps_['scale'])
            Pp = tuple(a.cpu().pin_memory() for a in h)

            def e2e1():
                so = bsl.compute_coefficients(Pp, sf, mf, verbose=0)
                return bsl.eval_particles(Pp, so.expcoef, sf, mf, verbose=0)
            for _ in range(3):
                e2e1()
            sync_all()
            t0 = time.perf_counter()
            for _ in range(10):
                e2e1()
            torch.cuda.synchronize()
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            entry['e2e'] = {'value': n * world * 10 / float(te.item()), 'unit': 'particles/s', 'ms_per_step': 1e2 * float(te.item()),
                            'h2d_bytes_per_step': (32 + 24) * n, 'd2h_bytes_per_step': 64 * n + 8 * 450,
                            'api': 'spheresl.compute_coefficients + spheresl.eval_particles, pinned host tensors in, NumPy arrays out '
                                   '(rank-local at N > 1)'}
        if cpu_ok:
            from oracle import oracle_np as O
            ps, ev, ef, xi, p0, d0 = sl
            ns = min(n, 100000)
            hx, hy, hz, hm = [a[:ns].cpu().numpy() for a in h]
            t0 = time.perf_counter()
            co = O.sl_accumulate(hx, hy, hz, hm, ps['lmax'], ps['nmax'], ev, ef, xi, p0, ps['cmap'], ps['scale'])
            fo = O.sl_all_eval_particles(hx, hy, hz, co, ps['lmax'], ps['nmax'], ev, ef, xi, p0, d0, ps['cmap'], ps['scale'])
            dt = time.perf_counter() - t0
            cg = H.accumulate(*[a[:ns] for a in h])
            H.contract(cg)
            fg = H.force(*[a[:ns] for a in h[:3]]).cpu().numpy()
            entry['cpu_baseline'] = {'value': ns / dt, 'unit': 'particles/s', 'cores': 1, 'kind': 'port',
                                     'sample': '%d particles, NumPy port of the reference formulation (oracle_np), one core, %.1f s' % (ns, dt)}
            entry['parity'] = {'coefficients_max_rel_err': _rel(cg.cpu().numpy(), co),
                               'force_max_rel_err': max(_rel(fg[i], fo[i]) for i in range(6)), 'sample': ns, 'tolerance': 1e-10}
        res['C1'] = entry
        del H, h

    # ---------------- C3 (+ C4): combined halo (SL lmax=6) + disc (EOF mmax=6) field
    if 'C3' in want or 'C4' in want:
        H6, sl6 = sl_handle(6)
        nb = max(min(int(5e6 * scale), int(1e8 * scale) // (2 * world)), 500)
        bd = dev_particles('disc', nb, 3003 + rank)
        bh = dev_particles('halo', nb, 3503 + rank)
        nc = min(nb, 1000000)
        c, s_ = E.accumulate(*[q[:nc] for q in bd])
        ch = H6.accumulate(*[q[:nc] for q in bh])
        cosc, sinc = c * 0.025, s_ * 0.025
        E.contract(cosc, sinc); H6.contract(ch)
    if 'C3' in want:
        n = max(int(1e8 * scale) // world, 1000)
        nd = n // 2

        def replicate(base, cnt):
            # distinct points with the base set's (R, z) distribution: the base rotated about z by a different angle per copy
            xs, ys, zs = [], [], []
            k = 0
            while sum(q.numel() for q in xs) < cnt:
                ca, sa = float(np.cos(0.61803 * k)), float(np.sin(0.61803 * k))
                xs.append(base[0] * ca - base[1] * sa); ys.append(base[0] * sa + base[1] * ca); zs.append(base[2]); k += 1
            return [torch.cat(v)[:cnt] for v in (xs, ys, zs)]
        pd_, ph_ = replicate(bd, nd), replicate(bh, n - nd)
        x = torch.cat([pd_[0], ph_[0]]); y = torch.cat([pd_[1], ph_[1]]); z = torch.cat([pd_[2], ph_[2]])
        del pd_, ph_
        box = {}

        torch.cuda.empty_cache()          # the earlier legs' cached blocks: the 6.4 GB result below should not need a cudaMalloc per call

        def c3():
            box.pop('out', None)          # the previous result goes back to the caching allocator first: one 6.4 GB block is re-used
            box['out'] = ops.field_force_cart(E, H6, x, y, z, rotpos=0.3)
        ms = timed(c3, 3, warm=1)
        kms = None
        ops.set_option('time_kernels', 1)
        c3()
        kms = ops.kernel_time_ms('field_sorted_pass')
        ops.set_option('time_kernels', 0)
        entry = {'config': 'C3 combined halo (SL lmax=6) + disc (EOF mmax=6) Cartesian force eval, 10^8 points (half disc-like, half halo-like)',
                 'particles': n * world, 'ms': ms, 'value': n * world / ms * 1e3, 'unit': 'particles/s',
                 'pbe_per_s': n * world / ms * 1e3 * (234 + 882), 'scaling': 'strong',
                 'tables': 'fp64', 'chunk': ops.get_option('field_sort_chunk'),
                 'roofline': _roof(FIELD_BYTES_PER_POINT * n, FIELD_FLOP_PER_POINT * n, ms, peaks, 'field_rec_kernel<6,6> (key-ordered pass)',
                                   note='pass = key + scan + scatter | evaluation | gather, two chunks in flight; evaluation kernel is bound by '
                                        'the L1 data pipe and FP64 issue (ncu, profiles/r02_summary.md)'),
                 'pass_ms_events_in_library': kms}
        out = box['out']
        idx = torch.cat([torch.arange(0, min(100000, nd), device=dev), torch.arange(nd, nd + min(100000, n - nd), device=dev)])
        # the key-ordered path against the caller-order kernel on a prefix of each half: same arithmetic, so bit-identical
        saved = ops.get_option('field_sort_min')
        ops.set_option('field_sort_min', 0)
        plain = ops.field_force_cart(E, H6, x[idx].contiguous(), y[idx].contiguous(), z[idx].contiguous(), rotpos=0.3)
        ops.set_option('field_sort_min', saved)
        entry['parity'] = {'key_ordered_equals_caller_order_bits': bool(torch.equal(out[:, idx], plain)), 'self_sample': int(idx.numel())}
        if cpu_ok:
            from oracle import oracle_np as O
            F = frozen(cosc.cpu().numpy(), sinc.cpu().numpy(), ch.cpu().numpy(), sl6)
            hx, hy, hz = x[idx].cpu().numpy(), y[idx].cpu().numpy(), z[idx].cpu().numpy()
            t0 = time.perf_counter()
            ref = O.fields_forces_cart(F, hx, hy, hz, rotpos=0.3)
            dt = time.perf_counter() - t0
            got = out[:, idx].cpu().numpy()
            entry['cpu_baseline'] = {'value': idx.numel() / dt, 'unit': 'particles/s', 'cores': 1, 'kind': 'port',
                                     'sample': '%d points (half disc, half halo), NumPy port of Fields.return_forces_cart '
                                               '(oracle_np), one core, %.1f s' % (idx.numel(), dt)}
            entry['parity'].update(vs_cpu_port_max_rel_err=max(_rel(got[i], ref[i]) for i in range(8)), sample=int(idx.numel()),
                                   tolerance=1e-10)
        res['C3'] = entry
        del x, y, z, out, box

    # ---------------- C4: 10^6 orbits x 10^4 leapfrog steps in the frozen field
    if 'C4' in want:
        norb = max(int(1e6 * scale) // world, 1000)
        nint = 10000 if scale >= 1.0 else max(int(10000 * scale), 100)
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
        box = {}
        ops.leapfrog(E, H6, P0, V0, min(nint, 200), 3e-4, rotfreq=-5.0)          # warm-up (workspace, tables in L2)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        box['st'] = ops.leapfrog(E, H6, P0, V0, nint, 3e-4, rotfreq=-5.0)[0]
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        st = box['st']
        steps_total = norb * world * (nint - 1)
        entry = {'config': 'C4 leapfrog orbit integration, 10^6 orbits x 10^4 steps in the frozen halo (SL lmax=6) + disc (EOF mmax=6) field, '
                           'dt=3e-4, rotfreq=-5', 'orbits': norb * world, 'steps': nint, 'ms': ms, 'value': steps_total / ms * 1e3,
                 'unit': 'orbit-steps/s', 'scaling': 'strong', 'tables': 'fp64', 'resort_every': ops.get_option('orbit_resort'),
                 'fraction_of_orbits_inside_R_lt_1': float((torch.sqrt(st[0] ** 2 + st[1] ** 2) < 1.0).double().mean().item()),
                 'roofline': _roof(96 * norb, (FIELD_FLOP_PER_POINT + 60) * norb * (nint - 1), ms, peaks,
                                   'leapfrog_perm_kernel<6,6> (key-ordered, re-sorted every %d steps)' % ops.get_option('orbit_resort'),
                                   note='state lives in registers / a 96-byte record per orbit: HBM side is 96 B per orbit per RUN')}
        # key-ordered path against the plain per-lane kernel: bit-identical end states
        ns_ = min(norb, 100000)
        k0 = ops.get_option('orbit_resort')
        a1 = ops.leapfrog(E, H6, P0[:, :ns_].contiguous(), V0[:, :ns_].contiguous(), 41, 3e-4, rotfreq=-5.0)[0]
        ops.set_option('orbit_resort', 0)
        a0 = ops.leapfrog(E, H6, P0[:, :ns_].contiguous(), V0[:, :ns_].contiguous(), 41, 3e-4, rotfreq=-5.0)[0]
        ops.set_option('orbit_resort', k0)
        entry['parity'] = {'key_ordered_equals_caller_order_bits': bool(torch.equal(a0, a1)), 'self_sample': '%d orbits x 40 steps' % ns_}
        if cpu_ok:
            from oracle import oracle_np as O
            F = frozen(cosc.cpu().numpy(), sinc.cpu().numpy(), ch.cpu().numpy(), sl6)
            no_, nst = min(norb, 2000), 101
            t0 = time.perf_counter()
            pr, vr, potr, _ = O.leapfrog(F, nst, 3e-4, pos0[:, :no_], vel0[:, :no_], rotfreq=-5.0)
            dt = time.perf_counter() - t0
            sg = ops.leapfrog(E, H6, pos0[:, :no_], vel0[:, :no_], nst, 3e-4, rotfreq=-5.0)[0].cpu().numpy()
            entry['cpu_baseline'] = {'value': no_ * (nst - 1) / dt, 'unit': 'orbit-steps/s', 'cores': 1, 'kind': 'port',
                                     'sample': '%d orbits x %d steps, NumPy port of integrate.leapfrog_integrate (oracle_np, '
                                               'vectorised over orbits), one core, %.1f s' % (no_, nst - 1, dt)}
            entry['parity'].update(vs_cpu_port_max_rel_err=max(max(_rel(sg[i], pr[i]), _rel(sg[3 + i], vr[i])) for i in range(3)),
                                   sample='%d orbits x %d steps' % (no_, nst - 1), tolerance=1e-8)
        res['C4'] = entry
        del P0, V0, box

    # ---------------- C5: coefficient time series, 200 snapshots x 10^7 particles (10^6 disc + 9x10^6 halo), one allreduce
    if 'C5' in want:
        H, sl = sl_handle(4)
        nsnap = 200 if scale >= 1.0 else max(int(200 * scale), 4)
        ndisc = max(int(1e6 * scale) // world, 1000)
        nhalo = max(int(9e6 * scale) // world, 1000)
        base_d = dev_particles('disc', ndisc, 5005 + rank)
        base_h = dev_particles('halo', nhalo, 5505 + rank)

        def snaps():
            for k in range(nsnap):      # distinct snapshots: the base set rotated by a different angle each time
                ca, sa = float(np.cos(0.01 * k)), float(np.sin(0.01 * k))
                yield ((base_d[0] * ca - base_d[1] * sa, base_d[0] * sa + base_d[1] * ca, base_d[2], base_d[3]),
                       (base_h[0] * ca - base_h[1] * sa, base_h[0] * sa + base_h[1] * ca, base_h[2], base_h[3]))
        box = {}

        def c5():
            box['out'] = parallel.accumulate_series(E, H, snaps())
        ms = timed(c5, 1, warm=1)
        npart = (ndisc + nhalo) * world * nsnap
        entry = {'config': 'C5 coefficient time series: 200 snapshots x 10^7 particles (10^6 disc EOF + 9x10^6 halo SL lmax=4) accumulation, '
                           'one allreduce of the [200, ncoef] block', 'snapshots': nsnap, 'particles_per_snapshot': (ndisc + nhalo) * world,
                 'ms': ms, 'value': npart / ms * 1e3, 'unit': 'particles/s', 'scaling': 'strong',
                 'roofline': _roof((32 + 24) * (ndisc + nhalo) * nsnap, None, ms, peaks, 'eof + sl sorted accumulate passes',
                                   note='32 B/particle accumulate read + 24 B/particle for the on-device rotation that makes each snapshot distinct')}
        cs, sn, ce = box['out']
        par = {}
        if world > 1:
            # the allreduced block against the sum formed in rank order from the gathered partials (snapshot 0)
            c0, s0 = E.accumulate(*next(snaps())[0])
            part = torch.cat([c0.reshape(-1), s0.reshape(-1)])
            gathered = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(gathered, part)
            want_ = gathered[0].clone()
            for r_ in range(1, world):
                want_ += gathered[r_]
            got_ = torch.cat([cs[0].reshape(-1), sn[0].reshape(-1)])
            par['allreduce_vs_rank_order_sum_max_rel_err'] = float(((got_ - want_).abs().max() / want_.abs().max()).item())
        if cpu_ok:
            from oracle import oracle_c as OC
            OC.use_all_cores()
            ps, ev, ef, xi, p0, d0 = sl
            nd_, nh_ = min(ndisc, 1000000), min(nhalo, 9000000)
            dx = [q[:nd_].cpu().numpy() for q in base_d]; hx = [q[:nh_].cpu().numpy() for q in base_h]
            t0 = time.perf_counter()
            cc, sc = OC.eof_accumulate(dx[0], dx[1], dx[2], dx[3], T['potC'], T['potS'], g)
            hc = OC.sl_accumulate(hx[0], hx[1], hx[2], hx[3], ps['lmax'], ps['nmax'], ev, ef, xi, p0, ps['cmap'], ps['scale'])
            dt = time.perf_counter() - t0
            cg, sg_ = E.accumulate(*[q[:nd_] for q in base_d])
            hg = H.accumulate(*[q[:nh_] for q in base_h])
            entry['cpu_baseline'] = {'value': (nd_ + nh_) / dt, 'unit': 'particles/s', 'cores': OC.threads(), 'kind': 'port',
                                     'sample': 'one snapshot of %d disc + %d halo particles, C port (oracle/bfe_oracle.c, OpenMP), %.1f s'
                                               % (nd_, nh_, dt)}
            par.update(eof_coefficients_max_rel_err=max(_rel(cg.cpu().numpy(), cc), _rel(sg_.cpu().numpy(), sc)),
                       sl_coefficients_max_rel_err=_rel(hg.cpu().numpy(), hc), sample=nd_ + nh_, tolerance=1e-10)
        entry['parity'] = par
        res['C5'] = entry
    return res


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from exptool_b200 import ops, synthetic as S, parallel
    from exptool_b200.basis import eof as beof

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the GPU arm has no CPU fallback)')
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # keep stdout to the one JSON line: libraries (NCCL's version banner) write to fd 1, so fd 1 is pointed at
        # stderr for the run and the JSON line goes to a private duplicate of the original stdout
        sys.stdout.flush()
        _json_fd = os.dup(1)
        os.dup2(2, 1)
        globals()['_JSON_OUT'] = os.fdopen(_json_fd, 'w')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    dev = torch.device('cuda', local)

    p, T, g = eof_setup()
    E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                      g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                      rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])

    # rotating particle sets: (32 B in + 48 B out) * 10^6 * NSETS > 2 x L2, so every step streams from HBM
    NSETS = 6
    sets = []
    for k in range(NSETS):
        x, y, z, m = S.exponential_disc(N_PART, 2002 + 100 * rank + k)
        sets.append(tuple(ops.dev(a) for a in (x, y, z, m)))
    outs = [torch.empty((6, N_PART), dtype=torch.float64, device=dev) for _ in range(NSETS)]
    lib = E.lib
    from exptool_b200.ops import _ptr, _stream
    from exptool_b200 import _lib as L
    import ctypes as C
    # Steps are independent particle sets (snapshots of a series), so NSTREAMS of them are kept in flight: each
    # stream has its own handle clone (shared device tables; own contraction, sorted-set workspace and counters)
    # and its own coefficient buffer.  A step is still prepare -> accumulate -> [allreduce] -> contract -> force
    # on ONE stream; the second stream fills the launch gaps, tails and latency-bound phases of the first
    # (profiles/dual_stream.py: 1 stream 174 us/step, 2 streams 138, 3 streams 133).
    NSTREAMS = max(1, args.streams)
    ops.set_option('grid_pct', args.grid_pct)
    if args.sort_stable >= 0:
        ops.set_option('sort_stable', args.sort_stable)
    for kv in args.opt:
        k_, v_ = kv.split('=')
        ops.set_option(k_, int(v_))
    Es = [E] + [E.clone() for _ in range(NSTREAMS - 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSTREAMS)]
    coefbufs = [torch.empty((2, g['mmax'] + 1, g['norder']), dtype=torch.float64, device=dev) for _ in range(NSTREAMS)]
    coefbuf = coefbufs[0]

    # ctypes arguments are built once (a c_void_p per pointer per call costs ~1 us of host time each)
    set_ptrs = [tuple(_ptr(t) for t in ts) for ts in sets]
    out_ptrs = [tuple(_ptr(o[i]) for i in range(6)) for o in outs]
    coef_ptrs = {id(cb): (_ptr(cb[0]), _ptr(cb[1])) for cb in coefbufs}
    stream_ptrs = {}
    mmax_, norder_ = g['mmax'], g['norder']

    # one NCCL communicator per stream (their allreduces then do not queue behind each other on one NCCL stream)
    pgs = {}
    if world > 1 and NSTREAMS > 1 and args.pg_per_stream:
        for st_ in streams:
            pgs[st_.cuda_stream] = dist.new_group(list(range(world)))

    # the per-step coefficient sum: one kernel over NVLink peer memory per stream (bfe_peer_allreduce), NCCL if
    # the peer buffers cannot be set up on this box
    peers = {}
    allreduce_kind = 'none'
    if world > 1:
        allreduce_kind = 'nccl'
        if args.allreduce == 'peer':
            try:
                for st_ in streams + [torch.cuda.current_stream()]:
                    peers[st_.cuda_stream] = parallel.PeerAllreduce(1024)
                allreduce_kind = 'peer-memory kernel (bfe_peer_allreduce)'
            except Exception as e_:
                peers = {}
                if rank == 0:
                    print('bench: peer-memory allreduce unavailable (%s); NCCL' % (e_,), file=sys.stderr)
        flag = torch.tensor([1.0 if peers else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)          # all ranks or none
        if float(flag.item()) == 0.0:
            peers = {}
            allreduce_kind = 'nccl'

    def step_on(k, Ei, coef, torch_stream):
        # one cell sort of the particle set serves both passes (include/bfe.h: bfe_eof_prepare)
        px, py, pz, pm = set_ptrs[k % NSETS]
        c0, c1 = coef_ptrs[id(coef)]
        st = stream_ptrs.get(torch_stream.cuda_stream)
        if st is None:
            st = stream_ptrs.setdefault(torch_stream.cuda_stream, C.c_void_p(torch_stream.cuda_stream))
        rc = lib.bfe_eof_prepare(Ei.h, N_PART, px, py, pz, pm, st)
        rc = rc or lib.bfe_eof_accumulate_prepared(Ei.h, c0, c1, st)
        if world > 1:
            pa = peers.get(torch_stream.cuda_stream)
            if pa is not None:
                rc = rc or lib.bfe_peer_allreduce(pa.h, c0, 2 * (mmax_ + 1) * norder_, st)
            else:
                with torch.cuda.stream(torch_stream):
                    dist.all_reduce(coef, group=pgs.get(torch_stream.cuda_stream))
        rc = rc or lib.bfe_eof_contract(Ei.h, c0, c1, 0, mmax_, norder_, 0, st)
        rc = rc or lib.bfe_eof_force_prepared(Ei.h, *out_ptrs[k % NSETS], st)
        if rc:
            L.check(rc)

    def step(k):                                   # single-stream step on the current stream (kernel timing legs)
        step_on(k, E, coefbuf, torch.cuda.current_stream())

    def run_steps(k0, nsteps):
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for k in range(k0, k0 + nsteps):
            i = k % NSTREAMS
            step_on(k, Es[i], coefbufs[i], streams[i])
        for st in streams:
            cur.wait_stream(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(0, args.warmup)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    th0 = time.perf_counter()
    run_steps(args.warmup, args.steps)
    host_issue_ms = 1e3 * (time.perf_counter() - th0) / args.steps      # CPU time to enqueue one step
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ops.launch_count() - launches0
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_total = float(tms.item())
    ms_per_step = ms_total / args.steps
    value = world * N_PART / (ms_per_step * 1e-3)

    # the same K steps one at a time on one stream: the latency of a step (and the throughput without overlap)
    for k in range(3):
        step(k)
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for k in range(args.steps):
        step(3 + k)
    l1.record()
    barrier()
    ms_single = l0.elapsed_time(l1) / args.steps

    # the price of bit-reproducibility: the same timed region with the slot claims of the cell sort done by global integer
    # atomics (option sort_stable = 0; results then agree to ~1e-16 from run to run instead of bit for bit)
    atomic_sort = None
    if ops.get_option('sort_stable') == 1:
        ops.set_option('sort_stable', 0)
        run_steps(0, max(args.warmup, 3))
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        run_steps(args.warmup, args.steps)
        a1.record()
        barrier()
        ops.set_option('sort_stable', 1)
        tat = torch.tensor([a0.elapsed_time(a1) / args.steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tat, op=dist.ReduceOp.MAX)
        atomic_sort = {'ms_per_step': float(tat.item()), 'value': world * N_PART / (float(tat.item()) * 1e-3),
                       'note': 'option sort_stable = 0: slot claims of the cell sort by global integer atomics, coefficients '
                               'reproducible to ~1e-16 only; the headline runs the stable atomic-free sort (bit-reproducible)'}

    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline object
    def time_kernel(fn, reps):
        for k in range(3):
            fn(k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(reps):
            fn(3 + k)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def k_prep(k):
        x, y, z, m = sets[k % NSETS]
        L.check(lib.bfe_eof_prepare(E.h, N_PART, _ptr(x), _ptr(y), _ptr(z), _ptr(m), _stream()))

    def k_acc(k):
        L.check(lib.bfe_eof_accumulate_prepared(E.h, _ptr(coefbuf[0]), _ptr(coefbuf[1]), _stream()))

    def k_con(k):
        L.check(lib.bfe_eof_contract(E.h, _ptr(coefbuf[0]), _ptr(coefbuf[1]), 0, g['mmax'], g['norder'], 0, _stream()))

    def k_force(k):
        o = outs[k % NSETS]
        L.check(lib.bfe_eof_force_prepared(E.h, *[_ptr(o[i]) for i in range(6)], _stream()))

    reps = max(min(args.steps, 50), 5)
    t_prep, t_acc = time_kernel(k_prep, reps), time_kernel(k_acc, reps)
    t_con, t_force = time_kernel(k_con, reps), time_kernel(k_force, reps)
    # live duration of every kernel of the step: CUDA events recorded inside the library around each launch
    # (bfe_set_option "time_kernels"), averaged over whole steps on rotating particle sets
    KNAMES = ['eof_cell_hist_kernel', 'eof_tile_colscan_kernel', 'eof_cell_scatter_kernel', 'eof_segsum_kernel', 'eof_node_contract_kernel',
              'eof_contract_kernel',
              'eof_force_sorted_kernel', 'eof_force_sorted_mma_kernel', 'eof_force_gather_kernel']
    ops.set_option('time_kernels', 1)
    ksum = {k: 0.0 for k in KNAMES}
    for k in range(reps):
        step(k)
        for nme in KNAMES:
            ksum[nme] += ops.kernel_time_ms(nme)
    ops.set_option('time_kernels', 0)
    kms = {k: v / reps for k, v in ksum.items() if v > 0.0}      # kernels that did not run report -1
    if sampler:
        sampler.stop()
    peak, peak_src = measured_peaks()
    # algorithmic HBM bytes per particle of the pass a kernel belongs to:
    # accumulate 32 B (x,y,z,m), force 72 B (x,y,z + six outputs); the contraction reads the six tables once.
    alg = {'eof_cell_hist_kernel': BYTES_ACC * N_PART, 'eof_tile_colscan_kernel': 2 * 148 * 128 * 64 * 4, 'eof_cell_scatter_kernel': BYTES_ACC * N_PART,
           'eof_segsum_kernel': BYTES_ACC * N_PART, 'eof_node_contract_kernel': 129 * 65 * 256 * 8,
           'eof_contract_kernel': 6 * 7 * 18 * 129 * 65 * 8,
           'eof_force_sorted_kernel': BYTES_FORCE * N_PART, 'eof_force_sorted_mma_kernel': BYTES_FORCE * N_PART,
           'eof_force_gather_kernel': BYTES_FORCE * N_PART}
    dom = max(kms, key=lambda k: kms[k])
    achieved = alg[dom] / (kms[dom] * 1e-3) / 1e9
    # dram bytes per launch from the ncu --set full capture of the same kernels (profiles/ncu_traffic.py; regenerated with the
    # kernels: r02 after the stable sort and the round-2 kernel changes, r01 as a fallback)
    traffic, traffic_all = None, None
    for tname in ('r02_traffic.json', 'r01_traffic.json'):
        tpath = os.path.join(ROOT, 'profiles', tname)
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj.get(dom)
            traffic_all = {k: tj[k] for k in kms if k in tj}
            break
    step_alg = (BYTES_ACC + BYTES_FORCE) * N_PART            # 104 B / particle for the whole step
    t_accpass = (kms['eof_cell_hist_kernel'] + kms.get('eof_tile_colscan_kernel', 0.0) + kms['eof_cell_scatter_kernel'] +
                 kms['eof_segsum_kernel'] + kms['eof_node_contract_kernel'])
    t_forcepass = sum(v for k, v in kms.items() if k.startswith('eof_force_'))
    # The executed FP64 work of the dominant kernel next to its HBM figure: per sorted record
    # 6 DMMA m8n8k4 per 8 records = 384 flop on the tensor pipe + ~57 flop per lane x 4 lanes = 228 flop of vector
    # FP64 (trig powers, harmonic sums, transposed reduce); nominal B200 FP64 peak 148 SMs x 64 lanes x 2 x 1.965 GHz.
    FP64_FLOP = {'eof_force_sorted_mma_kernel': 384 + 228}
    dfma_peak, dmma_peak = ops.fp64_peak('dfma'), ops.fp64_peak('dmma')          # measured on this device, in this run
    peaks = {'hbm_gbs': peak, 'hbm_source': peak_src, 'dfma_tflops': dfma_peak, 'dmma_tflops': dmma_peak}
    fp64 = None
    if dom in FP64_FLOP:
        tf = FP64_FLOP[dom] * N_PART / (kms[dom] * 1e-3) / 1e12
        fp64 = {'flop_per_particle': FP64_FLOP[dom], 'achieved_tflops': tf, 'peak_tflops': dfma_peak,
                'peak_source': 'measured in this run: bfe_fp64_peak (DFMA %.1f, DMMA %.1f TFLOP/s; nominal 148 SMs x 64 lanes x 2 x '
                               '1.965 GHz = 37.2)' % (dfma_peak, dmma_peak), 'frac': tf / dfma_peak}
    # bound = whichever of the two fractions is larger (SURVEY.md section 8d rule v); both sides are always in the object
    hbm_side = {'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'peak_source': peak_src}
    if fp64 is not None and fp64['frac'] > achieved / peak:
        top = {'bound': 'fp64', 'achieved': fp64['achieved_tflops'], 'peak': dfma_peak, 'unit': 'TFLOP/s', 'frac': fp64['frac']}
    else:
        top = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak}
    roofline = {'kernel': dom, 'fp64': fp64, 'hbm': hbm_side, 'traffic': traffic, 'peak_source': peak_src,
                'fp64_peaks_measured_tflops': {'dfma': dfma_peak, 'dmma': dmma_peak},
                'algorithmic_bytes_per_launch': alg[dom],
                'kernel_ms': kms,
                'api_call_ms': {'bfe_eof_prepare': t_prep, 'bfe_eof_accumulate_prepared': t_acc,
                                'bfe_eof_contract': t_con, 'bfe_eof_force_prepared': t_force},
                'passes': {'accumulate (hist+scatter+segsum+node_contract)': {'ms': t_accpass, 'algorithmic_bytes': BYTES_ACC * N_PART,
                                                                'achieved': BYTES_ACC * N_PART / (t_accpass * 1e-3) / 1e9,
                                                                'frac': BYTES_ACC * N_PART / (t_accpass * 1e-3) / 1e9 / peak},
                           'force (force_sorted+gather)': {'ms': t_forcepass, 'algorithmic_bytes': BYTES_FORCE * N_PART,
                                                           'achieved': BYTES_FORCE * N_PART / (t_forcepass * 1e-3) / 1e9,
                                                           'frac': BYTES_FORCE * N_PART / (t_forcepass * 1e-3) / 1e9 / peak}},
                'step': {'algorithmic_bytes': step_alg, 'achieved': step_alg / (ms_per_step * 1e-3) / 1e9,
                         'frac': step_alg / (ms_per_step * 1e-3) / 1e9 / peak,
                         'traffic': (sum(traffic_all.values()) if traffic_all and len(traffic_all) == len(kms) else None)},
                'traffic_per_kernel': traffic_all}
    roofline.update(top)

    # ---- e2e through the reference-facing API with host buffers (pinned), copies inside the timed region
    hx, hy, hz, hm = [torch.from_numpy(a).pin_memory() for a in S.exponential_disc(N_PART, 4004 + rank)]
    P = (hx, hy, hz, hm)
    geo = (g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'])
    tabs_acc = (T['potC'], T['potS'], g['mmax'], g['norder']) + geo + (g['ascale'], g['hscale'], g['cmap'])

    def e2e_step():
        if world == 1:
            c, s = beof.make_coefficients_multi(P, 1, *tabs_acc)
        else:
            # weak scaling: this rank's own 10^6 particles; partial coefficients summed with one allreduce
            Eh = beof.device_tables(*tabs_acc)
            c, s = parallel.eof_accumulate_host(Eh, hx, hy, hz, hm, already_sharded=True)
        return beof.accumulated_eval_particles(P, c, s, potC=T['potC'], rforceC=T['rforceC'], zforceC=T['zforceC'],
                                               potS=T['potS'], rforceS=T['rforceS'], zforceS=T['zforceS'],
                                               rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'],
                                               numy=g['numy'], MMAX=g['mmax'], NMAX=g['norder'], ASCALE=g['ascale'],
                                               HSCALE=g['hscale'], CMAP=g['cmap'], verbose=0)

    res = None
    for _ in range(max(args.warmup, 3)):
        res = e2e_step()                      # keep the previous result alive, as the timed loop does: the pinned
    barrier()                                 # result buffers (two alternate) then come from the allocator's cache
    t0 = time.perf_counter()
    per_call = []
    for _ in range(args.steps):
        tc = time.perf_counter()
        res = e2e_step()                      # returns NumPy arrays: each call ends with its own stream sync
        per_call.append(time.perf_counter() - tc)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    sys.stderr.write('e2e per-call ms: ' + ' '.join('%.2f' % (1e3 * t) for t in per_call) + '\n')
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N_PART * args.steps / float(te.item())
    reused = ops.get_option('host_reused_last') == 1
    e2e = {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': (32 if reused else 56) * N_PART,
           'd2h_bytes_per_step': 48 * N_PART + 2 * 8 * 126, 'ms_per_step': 1e3 * float(te.item()) / args.steps,
           'ms_per_step_median': 1e3 * float(np.median(per_call)),
           'one_upload_per_snapshot': reused,
           'api': ('eof.make_coefficients_multi' if world == 1 else 'parallel.eof_accumulate_host') +
                  ' + eof.accumulated_eval_particles, pinned host tensors in, NumPy arrays out; chunked '
                  'H2D | kernels | D2H pipeline on three streams inside each call; the accumulation uploads x,y,z,m every step '
                  '(32 B/particle), the evaluation of the same arrays runs from that copy (option host_reuse)'}
    # the same two calls with what a drop-in caller passes: plain (pageable) NumPy arrays
    Pp = tuple(np.array(t.numpy(), copy=True) for t in P)

    def e2e_step_pageable():
        if world == 1:
            c, s = beof.make_coefficients_multi(Pp, 1, *tabs_acc)
        else:
            Eh = beof.device_tables(*tabs_acc)
            c, s = parallel.eof_accumulate_host(Eh, Pp[0], Pp[1], Pp[2], Pp[3], already_sharded=True)
        return beof.accumulated_eval_particles(Pp, c, s, potC=T['potC'], rforceC=T['rforceC'], zforceC=T['zforceC'],
                                               potS=T['potS'], rforceS=T['rforceS'], zforceS=T['zforceS'],
                                               rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'],
                                               numy=g['numy'], MMAX=g['mmax'], NMAX=g['norder'], ASCALE=g['ascale'],
                                               HSCALE=g['hscale'], CMAP=g['cmap'], verbose=0)
    for _ in range(3):
        res = e2e_step_pageable()
    barrier()
    t0 = time.perf_counter()
    npg = max(3, min(args.steps, 10))
    for _ in range(npg):
        res = e2e_step_pageable()
    torch.cuda.synchronize()
    tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    e2e['pageable'] = {'value': world * N_PART * npg / float(tp.item()), 'unit': UNIT, 'ms_per_step': 1e3 * float(tp.item()) / npg,
                       'note': 'same calls with plain NumPy (pageable) inputs: what a drop-in caller passes'}

    # ---- numerical evidence for the multi-GPU sum: the allreduced coefficients of one step against the sum formed in
    #      rank order from the gathered partial blocks (bit-identical for the peer-memory kernel, which sums in rank order)
    parity = None
    if world > 1:
        px, py, pz, pm = sets[0]
        c0, s0 = E.accumulate(px, py, pz, pm)
        part = torch.stack([c0, s0]).contiguous()
        gathered = [torch.empty_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)
        want_ = gathered[0].clone()
        for r_ in range(1, world):
            want_ += gathered[r_]
        got_ = part.clone()
        pa = peers.get(torch.cuda.current_stream().cuda_stream)
        if pa is not None:
            pa.allreduce_(got_)
        else:
            dist.all_reduce(got_)
        err = float(((got_ - want_).abs().max() / want_.abs().max()).item())
        parity = {'allreduce_vs_rank_order_sum_max_rel_err': err, 'bit_identical': bool(torch.equal(got_, want_)),
                  'allreduce': allreduce_kind, 'tolerance': 1e-14}

    configs = None
    if args.configs != 'none':
        configs = run_configs(args, E, T, g, world, rank, dev, peaks)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline()

    if rank == 0:
        clocks = sampler.summary() if sampler else None
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': {'workload': WORKLOAD, 'particles_per_gpu': N_PART,
                           'l2': 'rotating %d particle sets (%.0f MB in+out) > 2x L2: every step streams from HBM'
                                 % (NSETS, NSETS * 80e6 / 1e6),
                           'streams': NSTREAMS,
                           'in_flight': '%d independent particle sets in flight on %d CUDA streams (cloned handles, '
                                        'shared tables); a step is serial on its stream' % (NSTREAMS, NSTREAMS),
                           'allreduce': allreduce_kind,
                           'parallelism': 'particles sharded over %d GPU(s), one 2 kB coefficient allreduce per step'
                                          % world if world > 1 else 'single GPU'},
                'host_issue_ms_per_step': host_issue_ms,
                'single_stream': {'ms_per_step': ms_single, 'value': world * N_PART / (ms_single * 1e-3)},
                'with_atomic_sort': atomic_sort,
                'pbe_per_s': value * PBE_PER_PARTICLE, 'e2e': e2e, 'gpu_launches': int(launches),
                'roofline': roofline, 'clocks': clocks}
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if parity is not None:
            line['parity'] = parity
        if configs is not None:
            line['configs'] = configs
        out = globals().get('_JSON_OUT') or sys.stdout
        out.write(json.dumps(line) + '\n')
        out.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--opt', action='append', default=[], help='library option name=value (repeatable; profiling A/B runs)')
    ap.add_argument('--streams', type=int, default=3, help='independent particle sets in flight (CUDA streams)')
    ap.add_argument('--allreduce', default='peer', choices=['peer', 'nccl'],
                    help='N>1: per-step coefficient sum by the peer-memory kernel (default) or NCCL')
    ap.add_argument('--grid-pct', type=int, default=100, help='share of the SMs the persistent grids of the step are sized for')
    ap.add_argument('--pg-per-stream', type=int, default=0, help='N>1: one NCCL communicator per stream (1) or shared (0)')
    ap.add_argument('--sort-stable', type=int, default=-1, help='option sort_stable of the cell sort (default: library default)')
    ap.add_argument('--configs', default='all', help="the other BASELINE configurations to run after the headline step: 'all', 'none' "
                                                      "or a comma list of C1,C3,C4,C5")
    ap.add_argument('--configs-scale', type=float, default=1.0, help='multiplies the particle / orbit / step counts of --configs')
    args = ap.parse_args()
    if args.allreduce == 'nccl':
        os.environ['BFE_PEER_ALLREDUCE'] = '0'          # the API-level (e2e) coefficient sum follows the same choice
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
