"""Would warp-UNIFORM tiles (every warp one (cell, interval) bin) pay?  2^15 disc points, each replicated 32 times by rotations
about z (same R, z, r => same cell and interval): after the key sort every bin holds a multiple of 32 points, so every warp of
the evaluation is uniform.  Compared with ordinary disc points, option stage_eval 0 / 1.
   python profiles/r02_uniform_warp_probe.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
import bench_configs as BC
E = BC.eof_handle(); H = BC.sl_handle(6)
n = 1 << 20
pd = BC.dev_particles('disc', n, 3003); ph = BC.dev_particles('halo', 1000000, 3503)
c, s = E.accumulate(*[p[:1000000] for p in pd]); ch = H.accumulate(*ph)
E.contract(c * 0.025, s * 0.025); H.contract(ch)
def timeit(fn, reps=8, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
res = {}
sets = {}
sets['ordinary_disc'] = (pd[0], pd[1], pd[2])
for rep in (32, 64, 256):
    nb = n // rep
    bx, by, bz = pd[0][:nb], pd[1][:nb], pd[2][:nb]
    ang = torch.arange(rep, device='cuda', dtype=torch.float64) * (2 * np.pi / rep)
    ca, sa = torch.cos(ang)[:, None], torch.sin(ang)[:, None]
    x = (bx[None, :] * ca - by[None, :] * sa).T.reshape(-1).contiguous()       # the replicas of a point are adjacent ...
    y = (bx[None, :] * sa + by[None, :] * ca).T.reshape(-1).contiguous()
    z = bz[None, :].expand(rep, nb).T.reshape(-1).contiguous()
    prm = torch.randperm(n, device='cuda')                                     # ... then shuffled: the sort has to find them
    sets['replicated_x%d' % rep] = (x[prm].contiguous(), y[prm].contiguous(), z[prm].contiguous())
ops.set_option('field_sort_chunk', 1 << 20)
for name, (x, y, z) in sets.items():
    ops.set_option('field_sort_min', 0)
    res[name + ' caller order'] = timeit(lambda: ops.field_force_cart(E, H, x, y, z, rotpos=0.3))
    ops.set_option('field_sort_min', 1)
    for st in (0, 1):
        ops.set_option('stage_eval', st)
        ops.set_option('time_kernels', 1)
        ops.field_force_cart(E, H, x, y, z, rotpos=0.3)
        ops.set_option('time_kernels', 0)
        res[name + ' key order stage_eval=%d' % st] = timeit(lambda: ops.field_force_cart(E, H, x, y, z, rotpos=0.3))
print(json.dumps(res, indent=1))
