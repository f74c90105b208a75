"""bench.py's e2e leg call by call: eof.make_coefficients_multi and eof.accumulated_eval_particles on pinned host
tensors, wall ms each, plus a cProfile of 20 steps.   python profiles/e2e_api_probe.py"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exptool_b200 import synthetic as S
from exptool_b200.basis import eof as beof
import bench
N = bench.N_PART
p, T, g = bench.eof_setup()
P = tuple(torch.from_numpy(a).pin_memory() for a in S.exponential_disc(N, 4004))
geo = (g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'])
tabs_acc = (T['potC'], T['potS'], g['mmax'], g['norder']) + geo + (g['ascale'], g['hscale'], g['cmap'])
kw = dict(potC=T['potC'], rforceC=T['rforceC'], zforceC=T['zforceC'], potS=T['potS'], rforceS=T['rforceS'],
          zforceS=T['zforceS'], rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'], numy=g['numy'],
          MMAX=g['mmax'], NMAX=g['norder'], ASCALE=g['ascale'], HSCALE=g['hscale'], CMAP=g['cmap'], verbose=0)
def step():
    c, s = beof.make_coefficients_multi(P, 1, *tabs_acc)
    return beof.accumulated_eval_particles(P, c, s, **kw)
for _ in range(5): res = step()
ta = tb = 0.0
for _ in range(20):
    t0 = time.perf_counter(); c, s = beof.make_coefficients_multi(P, 1, *tabs_acc); t1 = time.perf_counter()
    res = beof.accumulated_eval_particles(P, c, s, **kw); t2 = time.perf_counter()
    ta += t1 - t0; tb += t2 - t1
print('make_coefficients_multi %.3f ms   accumulated_eval_particles %.3f ms' % (ta / 20 * 1e3, tb / 20 * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(20): res = step()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
