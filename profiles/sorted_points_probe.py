"""Does cell-coherent ordering of the POINTS speed up the per-point field kernels?  Sort the points by (EOF cell, SL interval)
with torch (prototype: the sort itself is not timed here) and time field_force_cart / leapfrog on the sorted vs the caller's order.
   python profiles/sorted_points_probe.py [n]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
from helpers import sl_tables
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
ps, ev, ef, xi, p0, d0 = sl_tables(dict(sl_params=dict(lmax=6), kind='smooth', seed=0))
H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
x, y, z, m = [ops.dev(a) for a in S.exponential_disc(n, 2002)]
c, s = E.accumulate(x, y, z, m); ch = H.accumulate(*[ops.dev(a) for a in S.hernquist_halo(n, 1001)])
E.contract(c * 0.025, s * 0.025); H.contract(ch)

def keys(x, y, z):
    R = torch.sqrt(x * x + y * y + 1e-10)
    X, Y, ix, iy = ops.eof_return_bins(R, z, g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'])
    r3 = torch.sqrt(R * R + z * z)
    q = r3 / ps['scale']
    xi_ = (q - 1) / (q + 1)
    ib = torch.clamp(((xi_ - float(xi[0])) / float(xi[1] - xi[0])).floor().long(), 0, ps['numr'] - 2)
    return (ix * g['numy'] + iy) * 4096 + ib

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

res = {}
perm = torch.argsort(keys(x, y, z))
xs, ys, zs = x[perm].contiguous(), y[perm].contiguous(), z[perm].contiguous()
for f32 in (0, 1):
    ops.set_option('table_fp32', f32)
    res['field_cart_us_fp32tab%d' % f32] = dict(caller_order=timeit(lambda: ops.field_force_cart(E, H, x, y, z, rotpos=0.3)),
                                                cell_sorted=timeit(lambda: ops.field_force_cart(E, H, xs, ys, zs, rotpos=0.3)))
ops.set_option('table_fp32', 0)
a = ops.field_force_cart(E, H, x, y, z, rotpos=0.3)[:, perm]; b = ops.field_force_cart(E, H, xs, ys, zs, rotpos=0.3)
res['sorted_equals_unsorted'] = bool(torch.equal(a, b))
# leapfrog: orbits sorted by initial cell, K steps at a time with a re-sort between launches
norb = min(n, 200000); nint = 96
pos0 = torch.stack([x, y, z])[:, :norb].contiguous()
R = torch.sqrt(pos0[0] ** 2 + pos0[1] ** 2) + 1e-6
vel0 = torch.stack([-pos0[1] / R, pos0[0] / R, torch.zeros_like(R)]) * 1.5
t_plain = timeit(lambda: ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=0.0), reps=3)
res['leapfrog_ns_per_orbit_step_plain'] = t_plain * 1e3 / (norb * nint)
for K in (4, 8, 16, 32):
    def run():
        st = torch.cat([pos0, vel0])
        idx = torch.arange(norb, device=st.device)
        done = 0
        while done < nint - 1:
            pm = torch.argsort(keys(st[0], st[1], st[2]))
            st = st[:, pm].contiguous(); idx = idx[pm]
            k = min(K, nint - 1 - done)
            st, _, _ = ops.leapfrog(E, H, st[:3], st[3:], k + 1, 3e-4, rotfreq=0.0)
            done += k
        out = torch.empty_like(st); out[:, idx] = st
        return out
    t = timeit(run, reps=3)
    res['leapfrog_ns_per_orbit_step_resort_every_%d (torch sort included)' % K] = t * 1e3 / (norb * nint)
ref, _, _ = ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=0.0)
res['leapfrog_resorted_vs_plain_relerr'] = float((run() - ref).abs().max() / ref.abs().max())
print(json.dumps(dict(n=n, res=res), indent=1))
