import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, tempfile
from exptool_b200 import ops, synthetic as S
from oracle import oracle_np as O
import bench
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
ps, ev, ef = S.make_sl_tables(dict(lmax=6))
with tempfile.TemporaryDirectory() as tmp:
    mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
    A = np.genfromtxt(mf, comments='!', skip_header=5)
xi, r, p0, d0 = O.sl_init_table(A[:, 0], A[:, 1], A[:, 3], ps['numr'], ps['rmin'], ps['rmax'], ps['cmap'], ps['scale'])
H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
d = [ops.dev(a) for a in S.exponential_disc(1000000, 2002)]
h = [ops.dev(a) for a in S.hernquist_halo(1000000, 1001)]
ops.set_option('eof_accumulate_mode', 1); ops.set_option('eof_force_mode', 1); ops.set_option('sl_accumulate_mode', 1)
c, s = E.accumulate(*d); ch = H.accumulate(*h)
E.contract(c, s); H.contract(ch)
for _ in range(2):
    E.force(*d[:3]); H.force(*h[:3])
    ops.field_force_cart(E, H, *d[:3], rotpos=0.3)
torch.cuda.synchronize()
