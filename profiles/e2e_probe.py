"""Host-array entry points (bfe_eof_accumulate_host / bfe_eof_force_host) timed per pipeline chunk size:
   python profiles/e2e_probe.py   -> ms per call for host_chunk in {n (one shot), 500k, 250k, 125k, auto}"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exptool_b200 import ops, synthetic as S
import bench
N = bench.N_PART
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
P = [torch.from_numpy(a).pin_memory() for a in S.exponential_disc(N, 4004)]
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
c, s = E.accumulate_host(*P)
E.contract(c, s)
for chunk in (N, 500000, 250000, 125000, 0):
    ops.set_option('host_chunk', chunk)
    ta = t(lambda: E.accumulate_host(*P)); tf = t(lambda: E.force_host(*P[:3]))
    print(json.dumps({'host_chunk': chunk, 'accumulate_host_ms': round(ta, 3), 'force_host_ms': round(tf, 3),
                      'sum_ms': round(ta + tf, 3), 'particles_per_s': N / (ta + tf) * 1e3}), flush=True)
