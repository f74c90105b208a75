"""Run a few hot-path steps for ncu (python profiles/prof_step.py [eof|sl|field] [n])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
from oracle import oracle_np as O   # geometry helper only (profiling script, not product)
import bench

which = sys.argv[1] if len(sys.argv) > 1 else 'eof'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
x, y, z, m = [ops.dev(a) for a in S.exponential_disc(n, 2002)]
if which in ('sl', 'field'):
    lmax = 6 if which == 'field' else 4
    ps, ev, ef = S.make_sl_tables(dict(lmax=lmax))
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
        A = np.genfromtxt(mf, comments='!', skip_header=5)
    xi, r, p0, d0 = O.sl_init_table(A[:, 0], A[:, 1], A[:, 3], ps['numr'], ps['rmin'], ps['rmax'], ps['cmap'], ps['scale'])
    H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
    xh, yh, zh, mh = [ops.dev(a) for a in S.hernquist_halo(n, 1001)]
for _ in range(reps):
    if which == 'eof':
        c, s = E.accumulate(x, y, z, m)
        E.contract(c, s)
        out = E.force(x, y, z)
    elif which == 'sl':
        c = H.accumulate(xh, yh, zh, mh)
        H.contract(c)
        out = H.force(xh, yh, zh)
    else:
        c, s = E.accumulate(x, y, z, m * 0.025)
        ch = H.accumulate(xh, yh, zh, mh)
        E.contract(c, s); H.contract(ch)
        out = ops.field_force_cart(E, H, x, y, z, rotpos=0.3)
        st, tr, ns = ops.leapfrog(E, H, torch.stack([x, y, z])[:, :200000], torch.zeros(3, 200000, dtype=torch.float64, device='cuda'), 20, 3e-4, rotfreq=-5.0)
torch.cuda.synchronize()
print('done', which, n)
