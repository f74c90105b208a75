"""SL sorted accumulate: register-formulation deposit (sl_deposit_mode 2) against the shared-memory slab
kernel (default), 4e6 Hernquist particles, lmax 4 and 6; CUDA events.   python profiles/sl_probe.py [n]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from exptool_b200 import ops, synthetic as S
from helpers import sl_tables
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
res = {}
for lmax in (4, 6):
    meta = dict(sl_params=dict(lmax=lmax), kind='smooth', seed=0)
    ps, ev, ef, xi, p0, d0 = sl_tables(meta)
    H = ops.SLTables(ps['lmax'], ps['nmax'], ps['numr'], ps['cmap'], ps['scale'], ev, ef, xi, p0, d0)
    h = [ops.dev(a) for a in S.hernquist_halo(n, 4)]
    out = {}
    for mode in (1, 2):
        ops.set_option('sl_accumulate_mode', 2); ops.set_option('sl_deposit_mode', mode)
        for _ in range(3): c = H.accumulate(*h)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): c = H.accumulate(*h)
        b.record(); torch.cuda.synchronize()
        out[mode] = (a.elapsed_time(b) / 10 * 1e3, c.cpu().numpy())
    ops.set_option('sl_accumulate_mode', 1)
    cd = H.accumulate(*[t[:200000] for t in h]).cpu().numpy()
    ops.set_option('sl_accumulate_mode', 2)
    cs = H.accumulate(*[t[:200000] for t in h]).cpu().numpy()
    ops.set_option('sl_accumulate_mode', 0); ops.set_option('sl_deposit_mode', 0)
    res['lmax%d' % lmax] = dict(slab_us=out[1][0], register_us=out[2][0],
                                relerr_register_vs_slab=float(np.max(np.abs(out[2][1] - out[1][1])) / np.max(np.abs(out[1][1]))),
                                relerr_sorted_vs_direct_200k=float(np.max(np.abs(cs - cd)) / np.max(np.abs(cd))))
print(json.dumps(dict(n=n, res=res), indent=1))
