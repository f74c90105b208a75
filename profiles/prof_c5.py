"""One C5-like snapshot (EOF + SL lmax=6 accumulation of n particles) for an ncu launch list: python profiles/prof_c5.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import torch
from exptool_b200 import ops
import bench_configs as BC
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
E = BC.eof_handle(); H = BC.sl_handle(6)
pd = BC.dev_particles('disc', n, 3003); ph = BC.dev_particles('halo', n, 3503)
for _ in range(3):
    c, s = E.accumulate(*pd); ch = H.accumulate(*ph)
torch.cuda.synchronize()
print('done', n)
