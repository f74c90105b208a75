"""Per-op timing dashboard (CUDA events).  python profiles/bench_kernels.py [n_disc] [n_halo] [lmax]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
from oracle import oracle_np as O   # geometry helper only
import bench

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nh = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
lmax = int(sys.argv[3]) if len(sys.argv) > 3 else 6
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
ps, ev, ef = S.make_sl_tables(dict(lmax=lmax))
import tempfile
with tempfile.TemporaryDirectory() as tmp:
    mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
    A = np.genfromtxt(mf, comments='!', skip_header=5)
xi, r, p0, d0 = O.sl_init_table(A[:, 0], A[:, 1], A[:, 3], ps['numr'], ps['rmin'], ps['rmax'], ps['cmap'], ps['scale'])
H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
NS = 4
disc = [[ops.dev(a) for a in S.exponential_disc(nd, 2002 + k)] for k in range(NS)]
halo = [[ops.dev(a) for a in S.hernquist_halo(nh, 1001 + k)] for k in range(NS)]

def timeit(fn, reps=10):
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps): fn(k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us

res = {}
c, s = E.accumulate(*disc[0]); ch = H.accumulate(*halo[0])
E.contract(c * 0.025, s * 0.025); H.contract(ch)
for mode, name in ((1, 'direct'), (2, 'sorted')):
    ops.set_option('eof_accumulate_mode', mode); ops.set_option('eof_force_mode', mode)
    res['eof_accumulate_' + name] = timeit(lambda k: E.accumulate(*disc[k % NS]))
    res['eof_force_' + name] = timeit(lambda k: E.force(*disc[k % NS][:3]))
ops.set_option('eof_accumulate_mode', 0); ops.set_option('eof_force_mode', 0)
res['eof_prepare'] = timeit(lambda k: E.prepare(*disc[k % NS]))
res['eof_accumulate_prepared'] = timeit(lambda k: E.accumulate_prepared())
res['eof_force_prepared'] = timeit(lambda k: E.force_prepared())
res['eof_contract'] = timeit(lambda k: E.contract(c, s))
for mode, name in ((1, 'direct'), (2, 'sorted')):
    ops.set_option('sl_accumulate_mode', mode)
    res['sl_accumulate_' + name] = timeit(lambda k: H.accumulate(*halo[k % NS]))
ops.set_option('sl_accumulate_mode', 0)
res['sl_contract'] = timeit(lambda k: H.contract(ch))
res['sl_force'] = timeit(lambda k: H.force(*halo[k % NS][:3]))
res['field_cart_disc_points'] = timeit(lambda k: ops.field_force_cart(E, H, *disc[k % NS][:3], rotpos=0.3))
res['field_cart_halo_points'] = timeit(lambda k: ops.field_force_cart(E, H, *halo[k % NS][:3], rotpos=0.3))
norb, nint = min(nd, 200000), 50
pos0 = torch.stack(disc[0][:3])[:, :norb].contiguous()
x, y = pos0[0], pos0[1]
R = torch.sqrt(x * x + y * y) + 1e-6
vel0 = torch.stack([-y / R, x / R, torch.zeros_like(x)]) * 1.5
res['leapfrog_%dorb_x_%dsteps' % (norb, nint)] = timeit(lambda k: ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=-5.0), reps=3)
res['leapfrog_us_per_orbit_step'] = res['leapfrog_%dorb_x_%dsteps' % (norb, nint)] / (norb * (nint))
print(json.dumps(dict(n_disc=nd, n_halo=nh, lmax=lmax, us=res), indent=1))
