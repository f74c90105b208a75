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
    for staged in (0, 1):
        ops.set_option('staged_eval', staged)
        ops.set_option('eof_force_mode', 1)
        E.force(*d[:3]); H.force(*h[:3])
        E.force_eval_points(np.abs(d[0]) + 1e-3, d[2], d[1]); H.force_eval_points(np.abs(h[0]) + 1e-3, np.clip(h[2], -1, 1), h[1])
        ops.field_force_cart(E, H, *d[:3], rotpos=0.2); ops.field_force_cyl(E, H, *h[:3], rotpos=0.2)
    E.prepare(*d); E.accumulate_prepared(); E.force_prepared()
    pos0 = np.stack(d[:3])[:, :777]; vel0 = np.zeros_like(pos0); vel0[1] = 1.0
    ops.leapfrog(E, H, pos0, vel0, 12, 3e-4, rotfreq=-5.0, traj_stride=3, apse=True, ap_max=2)
    ops.leapfrog(E, H, pos0, vel0, 9, np.full(777, 2e-4), rotfreq=1.0, traj_stride=1)
torch.cuda.synchronize()
print('sanitize_run done')
