"""Live per-kernel durations of the bench step at 10^6 particles WITHOUT an L2 flush between kernels (the bench
condition: six rotating particle sets), CUDA events inside the library; plus the whole-step time with PDL.
   [BFE_LIB=variant.so] python profiles/kernel_times.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exptool_b200 import ops, synthetic as S, _lib as L
from exptool_b200.ops import _ptr, _stream
import bench
N = int(os.environ.get('AB_N', bench.N_PART))
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
NSETS = 6
sets = [tuple(ops.dev(a) for a in S.exponential_disc(N, 2002 + k)) for k in range(NSETS)]
outs = [torch.empty((6, N), dtype=torch.float64, device='cuda') for _ in range(NSETS)]
coef = torch.empty((2, g['mmax'] + 1, g['norder']), dtype=torch.float64, device='cuda')
lib = E.lib
def step(k):
    x, y, z, m = sets[k % NSETS]; o = outs[k % NSETS]
    L.check(lib.bfe_eof_prepare(E.h, N, _ptr(x), _ptr(y), _ptr(z), _ptr(m), _stream()))
    L.check(lib.bfe_eof_accumulate_prepared(E.h, _ptr(coef[0]), _ptr(coef[1]), _stream()))
    L.check(lib.bfe_eof_contract(E.h, _ptr(coef[0]), _ptr(coef[1]), 0, g['mmax'], g['norder'], 0, _stream()))
    L.check(lib.bfe_eof_force_prepared(E.h, *[_ptr(o[i]) for i in range(6)], _stream()))
def timed(steps=200):
    for k in range(10): step(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(steps): step(10 + k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3
us = min(timed(), timed(), timed())
KN = ['eof_cell_hist_kernel', 'eof_cell_scatter_kernel', 'eof_segsum_kernel', 'eof_node_contract_kernel',
      'eof_contract_kernel', 'eof_force_sorted_mma_kernel', 'eof_force_gather_kernel']
ops.set_option('time_kernels', 1)
acc = {k: 0.0 for k in KN}
reps = 50
for r in range(reps):
    step(r)
    for k in KN: acc[k] += max(ops.kernel_time_ms(k), 0.0)
ops.set_option('time_kernels', 0)
print(os.environ.get('BFE_LIB', 'default'), 'step us %.1f' % us, json.dumps({k.replace('eof_', '').replace('_kernel', ''): round(v / reps * 1e3, 1) for k, v in acc.items()}), flush=True)
