"""Per-warp cycle breakdown of sl_deposit_kernel (library built with BFE_NVCC_FLAGS=-DBFE_PROFILE_DEPOSIT)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, tempfile
from exptool_b200 import ops, synthetic as S, _lib
from oracle import oracle_np as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
lmax = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ps, ev, ef = S.make_sl_tables(dict(lmax=lmax))
with tempfile.TemporaryDirectory() as tmp:
    mf = S.write_hernquist_model(os.path.join(tmp, 'm'), a=ps['scale'])
    A = np.genfromtxt(mf, comments='!', skip_header=5)
xi, r, p0, d0 = O.sl_init_table(A[:, 0], A[:, 1], A[:, 3], ps['numr'], ps['rmin'], ps['rmax'], ps['cmap'], ps['scale'])
H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
h = [ops.dev(a) for a in S.hernquist_halo(n, 1001)]
lib = _lib.load()
dbg = torch.zeros((148 * 3 * 4, 8), dtype=torch.int64, device='cuda')
lib.bfe_sl_debug_set.argtypes = [C.c_void_p]
assert lib.bfe_sl_debug_set(C.c_void_p(dbg.data_ptr())) == 0
ops.set_option('sl_accumulate_mode', 2)
for _ in range(3):
    c = H.accumulate(*h)
torch.cuda.synchronize()
d = dbg.cpu().numpy(); d = d[d[:, 0] > 0]
names = ['total', 'wait_rec', 'expand', 'sum', 'flush', 'nflush', 'ntask', '-']
for i, nme in enumerate(names[:7]):
    v = d[:, i]
    print('%-9s min %8d med %8d mean %10.1f max %8d sum %12d' % (nme, v.min(), np.median(v), v.mean(), v.max(), v.sum()))
for i in np.argsort(d[:, 0])[-4:]:
    print('  warp', i, dict(zip(names, d[i])))
