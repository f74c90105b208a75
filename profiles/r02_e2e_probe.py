"""Where the e2e step's time goes (10^6 particles, pinned inputs): Python API calls vs ops-level vs C-level with sync.
   python profiles/r02_e2e_probe.py"""
import os, sys, time, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from exptool_b200 import ops, synthetic as S, _lib
from exptool_b200.basis import eof as beof
p, T, g = bench.eof_setup()
N = 1000000
hx, hy, hz, hm = [torch.from_numpy(a).pin_memory() for a in S.exponential_disc(N, 4004)]
P = (hx, hy, hz, hm)
geo = (g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'])
tabs_acc = (T['potC'], T['potS'], g['mmax'], g['norder']) + geo + (g['ascale'], g['hscale'], g['cmap'])
kw = dict(potC=T['potC'], rforceC=T['rforceC'], zforceC=T['zforceC'], potS=T['potS'], rforceS=T['rforceS'], zforceS=T['zforceS'],
          rmin=g['XMIN'], dR=g['dX'], zmin=g['YMIN'], dZ=g['dY'], numx=g['numx'], numy=g['numy'], MMAX=g['mmax'], NMAX=g['norder'],
          ASCALE=g['ascale'], HSCALE=g['hscale'], CMAP=g['cmap'], verbose=0)
def tm(fn, reps=20):
    for _ in range(4): r = fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
res = {}
c, s = beof.make_coefficients_multi(P, 1, *tabs_acc)
res['api_make_coefficients_multi_ms'] = tm(lambda: beof.make_coefficients_multi(P, 1, *tabs_acc))
res['api_accumulated_eval_particles_ms (after accumulate: reuse)'] = tm(lambda: beof.accumulated_eval_particles(P, c, s, **kw))
Ea = beof.device_tables(*tabs_acc)
Ef = beof.device_tables(T['potC'], T['potS'], g['mmax'], g['norder'], *geo, g['ascale'], g['hscale'], g['cmap'], rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
res['ops_accumulate_host_ms'] = tm(lambda: Ea.accumulate_host(hx, hy, hz, hm))
Ef.contract(c, s)
res['ops_force_host_ms (reuse)'] = tm(lambda: Ef.force_host(hx, hy, hz))
ops.set_option('host_reuse', 0)
res['ops_force_host_ms (no reuse)'] = tm(lambda: Ef.force_host(hx, hy, hz))
ops.set_option('host_reuse', 1); Ea.accumulate_host(hx, hy, hz, hm)
res['ops_contract_ms'] = tm(lambda: Ef.contract(c, s))
res['device_tables_lookup_ms'] = tm(lambda: beof.device_tables(*tabs_acc))
res['device_tables_lookup_force_ms'] = tm(lambda: beof.device_tables(T['potC'], T['potS'], g['mmax'], g['norder'], *geo, g['ascale'], g['hscale'], g['cmap'], rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS']))
res['pinned_empty_6n_ms'] = tm(lambda: ops.pinned_empty((6, N)))
lib = _lib.load(); st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
buf = torch.empty((2, 7, 18), dtype=torch.float64, device='cuda')
out = ops.pinned_empty((6, N))
pp = lambda t: C.c_void_p(t.data_ptr())
def c_acc():
    lib.bfe_eof_accumulate_host(Ea.h, N, pp(hx), pp(hy), pp(hz), pp(hm), pp(buf[0]), pp(buf[1]), st); torch.cuda.current_stream().synchronize()
def c_force():
    lib.bfe_eof_force_host(Ef.h, N, pp(hx), pp(hy), pp(hz), *[pp(out[i]) for i in range(6)], st); torch.cuda.current_stream().synchronize()
res['c_accumulate_host_sync_ms'] = tm(c_acc)
res['c_force_host_sync_ms (reuse)'] = tm(c_force)
for chunk in (125000, 250000, 500000, 1000000):
    ops.set_option('host_chunk', chunk)
    res['c_acc+force chunk %d ms' % chunk] = tm(lambda: (c_acc(), c_force()))
ops.set_option('host_chunk', 0)
# pageable inputs
px, py, pz, pm = [np.array(t.numpy(), copy=True) for t in P]
res['ops_accumulate_host_pageable_ms'] = tm(lambda: Ea.accumulate_host(px, py, pz, pm), reps=5)
res['ops_force_host_pageable_ms'] = tm(lambda: Ef.force_host(px, py, pz), reps=5)
print(json.dumps(res, indent=1))
