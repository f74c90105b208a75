"""SL sorted accumulation, stable tile sort + static deposit tasks against atomic claims + dynamic queue: python profiles/r02_sl_stable_probe.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import torch
from exptool_b200 import ops
import bench_configs as BC
res = {}
FC = [int(v) for v in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['32'])]
for lmax in (4, 6):
    H = BC.sl_handle(lmax)
    for n in (300000, 1000000, 4000000, 10000000):
        p = BC.dev_particles('halo', n, 77)
        for st in FC + [-1]:
            ops.set_option('sort_stable', 0 if st < 0 else 1)
            if st >= 0: ops.set_option('sl_flush_cost', st)
            for _ in range(3): H.accumulate(*p)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10): H.accumulate(*p)
            b.record(); torch.cuda.synchronize()
            res['lmax%d_n%d_%s_us_per_1e6' % (lmax, n, ('stable_fc%d' % st) if st >= 0 else 'atomic')] = a.elapsed_time(b) / 10 * 1e3 / n * 1e6
        del p
ops.set_option('sort_stable', 1)
print(json.dumps(res, indent=1))
if len(sys.argv) > 1: json.dump(res, open(sys.argv[1], 'w'), indent=1)
