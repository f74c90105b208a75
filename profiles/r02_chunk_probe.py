"""Key-ordered point evaluation on large halo-like / disc-like sets as a function of the chunk size and the key's low bits:
   python profiles/r02_chunk_probe.py [n=40000000] [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'profiles'))
import torch
from exptool_b200 import ops
import bench_configs as BC
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000000
SUBS = [int(v) for v in (sys.argv[3].split(',') if len(sys.argv) > 3 else ['4', '8'])]
CHUNKS = [int(v) for v in (sys.argv[4].split(',') if len(sys.argv) > 4 else ['1048576', '4194304', '16777216'])]
E = BC.eof_handle(); H = BC.sl_handle(6)
pd = BC.dev_particles('disc', 1000000, 3003); ph = BC.dev_particles('halo', 1000000, 3503)
c, s = E.accumulate(*pd); ch = H.accumulate(*ph)
E.contract(c * 0.025, s * 0.025); H.contract(ch)


def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


res = {}
ops.set_option('field_sort_min', 1)
for kind in ('halo', 'disc'):
    p = BC.dev_particles(kind, n, 5005)
    x, y, z = p[0], p[1], p[2]
    del p
    box = {}
    for sub in SUBS:
        ops.set_option('key_subbits', sub)
        for chunk in CHUNKS:
            ops.set_option('field_sort_chunk', chunk)
            def f():
                box.pop('o', None)
                box['o'] = ops.field_force_cart(E, H, x, y, z, rotpos=0.3)
            t = timeit(f) / n * 1e6
            res['%s_sub%d_chunk%d_us_per_1e6' % (kind, sub, chunk)] = t
            print(kind, sub, chunk, round(t, 1), flush=True)
    del x, y, z, box
    torch.cuda.empty_cache()
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], 'w'), indent=1)
