import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import torch
from exptool_b200 import ops
import bench_configs as BC
H = BC.sl_handle(6)
p = BC.dev_particles('halo', 1000000, 77)
for st in (1, 0, 1, 0):
    ops.set_option('sort_stable', st)
    H.accumulate(*p)
torch.cuda.synchronize()
