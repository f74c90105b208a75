"""A/B of the per-lane block evaluation with 256-bit loads (option blk_eval) against the default kernels:
field_force_cart on 10^6 disc / halo points and leapfrog (orbits x steps), plus a result comparison.
   python profiles/blk_ab.py [n] [norbit] [nint]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
from oracle import oracle_np as O   # geometry helper only
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
norb = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
nint = int(sys.argv[3]) if len(sys.argv) > 3 else 100
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
ps, ev, ef = S.make_sl_tables(dict(lmax=6))
import tempfile
with tempfile.TemporaryDirectory() as tmp:
    mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
    A = np.genfromtxt(mf, comments='!', skip_header=5)
xi, r, p0, d0 = O.sl_init_table(A[:, 0], A[:, 1], A[:, 3], ps['numr'], ps['rmin'], ps['rmax'], ps['cmap'], ps['scale'])
H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
NS = 4
disc = [[ops.dev(a) for a in S.exponential_disc(n, 2002 + k)] for k in range(NS)]
halo = [[ops.dev(a) for a in S.hernquist_halo(n, 1001 + k)] for k in range(NS)]
c, s = E.accumulate(*disc[0]); ch = H.accumulate(*halo[0])
E.contract(c * 0.025, s * 0.025); H.contract(ch)

def timeit(fn, reps=10):
    for k in range(3): fn(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps): fn(k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us

pos0 = torch.stack(disc[0][:3])[:, :norb].contiguous()
x, y = pos0[0], pos0[1]
R = torch.sqrt(x * x + y * y) + 1e-6
vel0 = torch.stack([-y / R, x / R, torch.zeros_like(x)]) * 1.5
res, outs = {}, {}
for name, staged, blk, f32 in (('per_lane', 0, 0, 0), ('staged', 1, 0, 0), ('blk256', 1, 1, 0), ('blk256_fp32_tables', 1, 1, 1)):
    ops.set_option('staged_eval', staged); ops.set_option('blk_eval', blk); ops.set_option('table_fp32', f32)
    r_ = {}
    r_['field_cart_disc_us'] = timeit(lambda k: ops.field_force_cart(E, H, *disc[k % NS][:3], rotpos=0.3))
    r_['field_cart_halo_us'] = timeit(lambda k: ops.field_force_cart(E, H, *halo[k % NS][:3], rotpos=0.3))
    r_['field_cyl_disc_us'] = timeit(lambda k: ops.field_force_cyl(E, H, *disc[k % NS][:3], rotpos=0.3))
    ops.set_option('eof_force_mode', 1)                     # unsorted points: the per-point kernels
    r_['eof_force_unsorted_us'] = timeit(lambda k: E.force(*disc[k % NS][:3]))
    ops.set_option('eof_force_mode', 0)
    r_['sl_force_us'] = timeit(lambda k: H.force(*halo[k % NS][:3]))
    t = timeit(lambda k: ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=-5.0), reps=3)
    r_['leapfrog_ns_per_orbit_step'] = t * 1e3 / (norb * nint)
    res[name] = r_
    st, tr, ns = ops.leapfrog(E, H, pos0, vel0, nint, 3e-4, rotfreq=-5.0)
    outs[name] = (ops.field_force_cart(E, H, *disc[1][:3], rotpos=0.3).cpu().numpy(),
                  ops.field_force_cart(E, H, *halo[1][:3], rotpos=0.3).cpu().numpy(), st.cpu().numpy())
ops.set_option('staged_eval', 1); ops.set_option('blk_eval', 1); ops.set_option('table_fp32', 0)
def rel(a, b): return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
res['blk_vs_per_lane_relerr'] = [rel(outs['blk256'][i], outs['per_lane'][i]) for i in range(3)]
res['fp32_tables_vs_fp64_relerr'] = [rel(outs['blk256_fp32_tables'][i], outs['per_lane'][i]) for i in range(3)]
print(json.dumps(dict(n=n, norbit=norb, nint=nint, res=res), indent=1))
