"""A few key-ordered point evaluations for ncu: python profiles/prof_field_split.py [split=1] [kind=disc] [n=1048576]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import torch
from exptool_b200 import ops
import bench_configs as BC
split = int(sys.argv[1]) if len(sys.argv) > 1 else 1
kind = sys.argv[2] if len(sys.argv) > 2 else 'disc'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
if split:
    ops.set_option('field_split', split)      # option of the removed two-kernel experiment (see bfe_orbit_sort.cu)
for kv in sys.argv[4:]:
    k_, v_ = kv.split('='); ops.set_option(k_, int(v_))
E = BC.eof_handle(); H = BC.sl_handle(6)
pd = BC.dev_particles('disc', 1000000, 3003); ph = BC.dev_particles('halo', 1000000, 3503)
c, s = E.accumulate(*pd); ch = H.accumulate(*ph)
E.contract(c * 0.025, s * 0.025); H.contract(ch)
p = BC.dev_particles(kind, n, 4004)
ops.set_option('field_sort_min', 1)
for _ in range(3):
    out = ops.field_force_cart(E, H, p[0], p[1], p[2], rotpos=0.3)
torch.cuda.synchronize()
print('done', split, kind, n)
