"""Per-warp cycle breakdown of eof_deposit_kernel.  Needs a library built with
   BFE_NVCC_FLAGS=-DBFE_PROFILE_DEPOSIT python exptool_b200/csrc/build.py --force"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S, _lib
import bench
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'])
x, y, z, m = [ops.dev(a) for a in S.exponential_disc(1000000, 2002)]
lib = _lib.load()
dbg = torch.zeros((296 * 8, 8), dtype=torch.int64, device='cuda')
lib.bfe_debug_set.argtypes = [C.c_void_p]
assert lib.bfe_debug_set(C.c_void_p(dbg.data_ptr())) == 0
E.prepare(x, y, z, m)
for _ in range(3):
    E.accumulate_prepared()
torch.cuda.synchronize()
d = dbg.cpu().numpy()
names = ['main', 'total', 'wait_rec', 'expand', 'mma', 'flush', 'nflush', 'nksteps']
for i, nme in enumerate(names):
    v = d[:, i]
    print('%-9s min %8d med %8d mean %10.1f max %8d sum %12d' % (nme, v.min(), np.median(v), v.mean(), v.max(), v.sum()))
print('slowest warps:')
for i in np.argsort(d[:, 0])[-5:]:
    print('  warp', i, dict(zip(names, d[i])))
