"""Cell-sorted leapfrog (bfe_orbit_sort.cu) against the plain kernel: ns per orbit-step for re-sort intervals K, end states
compared bit for bit.   python profiles/orbit_sort_probe.py [norbit] [nint]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
from helpers import sl_tables
import bench
norb = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nint = int(sys.argv[2]) if len(sys.argv) > 2 else 257
kind = sys.argv[3] if len(sys.argv) > 3 else 'disc'        # initial conditions: disc particles or halo particles
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
ps, ev, ef, xi, p0, d0 = sl_tables(dict(sl_params=dict(lmax=6), kind='smooth', seed=0))
H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
x, y, z, m = [ops.dev(a) for a in S.exponential_disc(norb, 4004)]
c, s = E.accumulate(x, y, z, m); ch = H.accumulate(*[ops.dev(a) for a in S.hernquist_halo(norb, 1001)])
E.contract(c * 0.025, s * 0.025); H.contract(ch)
if kind == 'halo':
    x, y, z, m = [ops.dev(a) for a in S.hernquist_halo(norb, 4005)]
pos0 = torch.stack([x, y, z]).contiguous()
a = ops.field_force_cart(E, H, x, y, z)
R = torch.sqrt(x * x + y * y) + 1e-12
fr = ((a[0] + a[1]) * x + (a[2] + a[3]) * y) / R
vc = torch.sqrt(torch.clamp(-R * fr, min=1e-12))
gen = torch.Generator(device='cuda'); gen.manual_seed(44)
f = 0.6 + 0.5 * torch.rand(norb, dtype=torch.float64, device='cuda', generator=gen)
vel0 = torch.stack([-y / R * vc * f, x / R * vc * f, 0.1 * vc * torch.randn(norb, dtype=torch.float64, device='cuda', generator=gen)]).contiguous()

def run():
    return ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=-5.0)

def timeit(reps=2):
    run(); torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(reps): out = run()
    b_.record(); torch.cuda.synchronize()
    return a_.elapsed_time(b_) / reps * 1e6 / (norb * (nint - 1)), out

res = {}
for f32 in (0, 1):
    ops.set_option('table_fp32', f32)
    ops.set_option('orbit_resort', 0)
    t, ref = timeit()
    res['fp32tab%d_plain' % f32] = t
    for K in (4, 8, 16, 32, 64):
        ops.set_option('orbit_resort', K)
        t, out = timeit()
        res['fp32tab%d_resort_%d' % (f32, K)] = t
        res['fp32tab%d_resort_%d_bit_identical' % (f32, K)] = bool(torch.equal(out[0], ref[0]) and torch.equal(out[2], ref[2]))
ops.set_option('table_fp32', 0); ops.set_option('orbit_resort', 16)
print(json.dumps(dict(norbit=norb, nint=nint, kind=kind, ns_per_orbit_step=res), indent=1))
