"""Fields.return_forces_cart on n points: caller's order (field_sort_min = 0) vs the cell-ordered path (bfe_orbit_sort.cu),
FP64 and FP32 tables, interleaved A/B/A/B so that clock drift shows.   python profiles/field_sort_probe.py [n]"""
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
sets = {'disc': [ops.dev(a) for a in S.exponential_disc(n, 2002)][:3], 'halo': [ops.dev(a) for a in S.hernquist_halo(n, 1001)][:3]}
xd, yd, zd = sets['disc']
c, s = E.accumulate(xd, yd, zd, torch.full_like(xd, 1.0 / n)); ch = H.accumulate(*sets['halo'], torch.full_like(xd, 1.0 / n))
E.contract(c * 0.025, s * 0.025); H.contract(ch)

def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
res = {}
for name, (x, y, z) in sets.items():
    for f32 in (0, 1):
        ops.set_option('table_fp32', f32)
        row = []
        for rep in range(2):
            for smin in (0, 1):
                ops.set_option('field_sort_min', smin)
                row.append(round(t(lambda: ops.field_force_cart(E, H, x, y, z, rotpos=0.3)), 1))
        res['%s_points_fp32tab%d [caller, sorted, caller, sorted] us' % (name, f32)] = row
ops.set_option('table_fp32', 0); ops.set_option('field_sort_min', 0)
print(json.dumps(dict(n=n, res=res), indent=1))
