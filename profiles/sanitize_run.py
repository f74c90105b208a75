"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
from helpers import eof_tables, sl_tables
for lmax in (4, 6):
    meta = dict(eof_params={}, sl_params=dict(lmax=lmax), kind='smooth', seed=0)
    p, T, g = eof_tables(meta)
    E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'],
                      g['ascale'], g['hscale'], g['cmap'], rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
    ps, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
    n = 5003
    d = S.exponential_disc(n, 3); h = S.hernquist_halo(n, 4)
    for mode in (1, 2):
        ops.set_option('eof_accumulate_mode', mode); ops.set_option('eof_force_mode', mode); ops.set_option('sl_accumulate_mode', mode)
        c, s = E.accumulate(*d); ch = H.accumulate(*h)
        E.contract(c, s); H.contract(ch)
        E.force(*d[:3]); H.force(*h[:3])
    pos0 = np.stack(d[:3])[:, :777]; vel0 = np.zeros_like(pos0); vel0[1] = 1.0
    # per-point kernel variants: per-lane rows, warp-staged, per-lane blocks with 256-bit loads, FP32 tables
    for staged, blk, f32 in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 1, 0), (1, 1, 1)):
        ops.set_option('staged_eval', staged); ops.set_option('blk_eval', blk); ops.set_option('table_fp32', f32)
        ops.set_option('eof_force_mode', 1)
        E.force(*d[:3]); H.force(*h[:3])
        E.force_eval_points(np.abs(d[0]) + 1e-3, d[2], d[1]); H.force_eval_points(np.abs(h[0]) + 1e-3, np.clip(h[2], -1, 1), h[1])
        ops.field_force_cart(E, H, *d[:3], rotpos=0.2); ops.field_force_cyl(E, H, *h[:3], rotpos=0.2)
        ops.leapfrog(E, H, pos0, vel0, 12, 3e-4, rotfreq=-5.0, traj_stride=3, apse=True, ap_max=2)
        ops.leapfrog(E, H, pos0, vel0, 9, np.full(777, 2e-4), rotfreq=1.0, traj_stride=1)
    ops.set_option('staged_eval', 1); ops.set_option('blk_eval', 1); ops.set_option('table_fp32', 0); ops.set_option('eof_force_mode', 0)
    # round 2: key-ordered field evaluation (two chunks in flight on two streams) and leapfrog, global loads / TMA-staged
    # shared-memory blocks, FP64 / FP32 tables, uneven chunks, re-sort every 1 / 3 steps, per-orbit step sizes
    saved = {k: ops.get_option(k) for k in ('field_sort_min', 'field_sort_chunk', 'orbit_sort_min', 'orbit_resort', 'stage_eval')}
    xm = np.concatenate([d[0], h[0]]); ym = np.concatenate([d[1], h[1]]); zm = np.concatenate([d[2], h[2]])
    ops.set_option('field_sort_min', 1); ops.set_option('field_sort_chunk', 3001); ops.set_option('orbit_sort_min', 1)
    for stage, f32 in ((0, 0), (1, 0), (0, 1)):
        ops.set_option('stage_eval', stage); ops.set_option('table_fp32', f32)
        ops.field_force_cart(E, H, xm, ym, zm, rotpos=0.2); ops.field_force_cyl(E, H, xm, ym, zm, rotpos=0.2)
        for K in (1, 3):
            ops.set_option('orbit_resort', K)
            ops.leapfrog(E, H, pos0, vel0, 12, 3e-4, rotfreq=-5.0)
            ops.leapfrog(E, H, pos0, vel0, 11, np.full(777, 2e-4), rotfreq=1.0)
    ops.set_option('table_fp32', 0)
    for k, v in saved.items():
        ops.set_option(k, v)
    # host-array pipelines: one upload per snapshot, pageable inputs through the copy threads, the SL pipeline
    E.accumulate_host(*d); E.force_host(*d[:3]); H.accumulate_host(*h); H.force_host(*h[:3])
    pin = [torch.from_numpy(a).pin_memory() for a in d]
    E.accumulate_host(*pin); E.force_host(*pin[:3])
    E.prepare(*d); E.accumulate_prepared(); E.force_prepared()
    for st in (1, 0):                         # stable tile sorts (EOF cells, SL radial bins, static deposit tasks) / atomic slot claims
        ops.set_option('sort_stable', st); ops.set_option('eof_accumulate_mode', 2); ops.set_option('sl_accumulate_mode', 2)
        E.accumulate(*d); H.accumulate(*h)
    ops.set_option('sort_stable', 1); ops.set_option('eof_accumulate_mode', 0); ops.set_option('sl_accumulate_mode', 0)
    # density outputs, building blocks
    H.contract_density(ch); H.density(*h[:3]); H.density_eval_points(np.abs(h[0]) + 1e-3, np.clip(h[2], -1, 1), h[1])
    H.radial_matrices(np.abs(h[0]) + 1e-6); ops.legendre_tables(lmax, np.clip(h[2], -1, 1))
    E.get_pot(np.abs(d[0]) + 1e-3, d[2])
    ops.eof_return_bins(np.abs(d[0]) + 1e-3, d[2], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'])
# the peer-memory sum with two emulated ranks on this GPU
import ctypes as C
from exptool_b200 import _lib
lib = _lib.load()
bufs = (C.c_void_p * 2)()
for r in range(2):
    pp_ = C.c_void_p(); hd = C.create_string_buffer(64)
    _lib.check(lib.bfe_peer_buffer_create(300, C.byref(pp_), hd)); bufs[r] = pp_.value
prs = []
for r in range(2):
    hh = C.c_void_p(); _lib.check(lib.bfe_peer_create(r, 2, 300, bufs, C.byref(hh))); prs.append(hh)
dat = torch.randn(2, 8, 252, dtype=torch.float64, device='cuda')
sts = [torch.cuda.Stream() for _ in range(2)]
torch.cuda.synchronize()
for it in range(8):
    for r in range(2):
        _lib.check(lib.bfe_peer_allreduce(prs[r], C.c_void_p(dat[r, it].data_ptr()), 252, C.c_void_p(sts[r].cuda_stream)))
torch.cuda.synchronize()
assert torch.equal(dat[0], dat[1])
torch.cuda.synchronize()
print('fp64 peaks', ops.fp64_peak('dfma'), ops.fp64_peak('dmma'))
print('sanitize_run done')
