import sys; import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import torch
from exptool_b200 import ops
import bench_configs as BC
E = BC.eof_handle(); H = BC.sl_handle(6)
pd = BC.dev_particles('disc', 1000000, 3003); ph = BC.dev_particles('halo', 1000000, 3503)
c, s = E.accumulate(*pd); ch = H.accumulate(*ph)
E.contract(c*0.025, s*0.025); H.contract(ch)
ops.set_option('field_sort_min', 1)
ops.field_force_cart(E, H, pd[0], pd[1], pd[2])
print('keycell_nkeys', ops.get_option('keycell_nkeys'))
