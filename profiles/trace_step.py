"""Device-side timeline of the prepared EOF step (needs BFE_NVCC_FLAGS=-DBFE_TRACE build):
   per kernel and phase: first / median / last timestamp over CTAs, relative to the step's first event (us)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S, _lib
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'],
                  g['ascale'], g['hscale'], g['cmap'], rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
x, y, z, m = [ops.dev(a) for a in S.exponential_disc(n, 2002)]
lib = _lib.load()
CAP = 200000
buf = torch.zeros(2 + 2 * CAP, dtype=torch.int64, device='cuda')
lib.bfe_debug_set_trace.argtypes = [C.c_void_p, C.c_uint]
def step():
    E.prepare(x, y, z, m); c, s = E.accumulate_prepared(); E.contract(c, s); E.force_prepared()
for _ in range(3):
    step()
torch.cuda.synchronize()
assert lib.bfe_debug_set_trace(C.c_void_p(buf.data_ptr()), CAP) == 0
for _ in range(3):       # three back-to-back steps; the middle one is analysed
    step()
torch.cuda.synchronize()
b = buf.cpu().numpy().astype(np.uint64)
cnt = int(b[0]); rec = b[2:2 + 2 * cnt].reshape(-1, 2)
kid = (rec[:, 0] >> np.uint64(56)).astype(int); ph = ((rec[:, 0] >> np.uint64(48)) & np.uint64(0xff)).astype(int)
t = rec[:, 1].astype(np.int64)
names = {0: 'hist', 1: 'scatter', 2: 'segsum', 3: 'node_contract', 5: 'force_mma', 6: 'gather'}
# split the three steps at the hist kernel starts
h0 = np.sort(t[(kid == 0) & (ph == 0)])
gaps = np.where(np.diff(h0) > 20000)[0]
starts = [h0[0]] + [h0[i + 1] for i in gaps]
print('events', cnt, 'steps found', len(starts), 'step period us', np.diff(starts) / 1e3)
lo, hi = starts[1], starts[2]
sel = (t >= lo) & (t < hi)
for k in sorted(names):
    for p_ in sorted(set(ph[sel & (kid == k)])):
        v = (t[sel & (kid == k) & (ph == p_)] - lo) / 1e3
        print('%-14s phase %d  n %5d  first %8.2f  median %8.2f  last %8.2f' % (names[k], p_, len(v), v.min(), np.median(v), v.max()))
