"""The benchmark step (prepare -> accumulate_prepared -> contract -> force_prepared) a few times, for ncu:
   ncu --set full --clock-control none --import-source on -k regex:'eof_cell|eof_deposit|eof_contract|eof_force|eof_seg|eof_node' \
       -s 12 -c 6 -o gpurun_out/step python profiles/prof_prepared.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exptool_b200 import ops, synthetic as S
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
x, y, z, m = [ops.dev(a) for a in S.exponential_disc(n, 2002)]
for _ in range(reps):
    E.prepare(x, y, z, m)
    c, s = E.accumulate_prepared()
    E.contract(c, s)
    out = E.force_prepared()
torch.cuda.synchronize()
print('done', n)
