"""Per-kernel live durations (CUDA events inside the library, bfe_kernel_time_ms) of the prepared EOF step as a
function of the particle count: the intercept at n -> 0 is each kernel's fixed cost (launch, zeroing, merges,
last-CTA epilogues), the slope its per-particle cost.   python profiles/kernel_scaling.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exptool_b200 import ops, synthetic as S
import bench

KN = ['eof_cell_hist_kernel', 'eof_cell_scatter_kernel', 'eof_segsum_kernel', 'eof_node_contract_kernel',
      'eof_contract_kernel', 'eof_force_sorted_mma_kernel', 'eof_force_sorted_kernel', 'eof_force_gather_kernel']
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
full = [ops.dev(a) for a in S.exponential_disc(4000000, 2002)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
out = {}
for n in (1000, 62500, 250000, 1000000, 4000000):
    x, y, z, m = [a[:n] for a in full]
    for _ in range(3):
        E.prepare(x, y, z, m); c, s = E.accumulate_prepared(); E.contract(c, s); E.force_prepared()
    ops.set_option('time_kernels', 1)
    acc = {k: 0.0 for k in KN}
    reps = 20
    for _ in range(reps):
        flush.zero_()                                   # L2 flush between repetitions
        E.prepare(x, y, z, m); c, s = E.accumulate_prepared(); E.contract(c, s); E.force_prepared()
        for k in KN:
            v = ops.kernel_time_ms(k)
            if v > 0:
                acc[k] += v
    ops.set_option('time_kernels', 0)
    out[n] = {k: round(v / reps * 1e3, 2) for k, v in acc.items() if v > 0}
    print(n, json.dumps(out[n]), flush=True)
