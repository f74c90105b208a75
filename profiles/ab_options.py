"""A/B of library options on the bench step (prepare -> accumulate_prepared -> contract -> force_prepared on six
rotating 10^6-particle sets, CUDA events):   python profiles/ab_options.py name=v[,v...] [name=v,...] ...
Every combination of the listed option values is timed, e.g.  pdl=0,1 l2_persist=0,1"""
import sys, os, json, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exptool_b200 import ops, synthetic as S, _lib as L
from exptool_b200.ops import _ptr, _stream
import bench

N = int(os.environ.get('AB_N', bench.N_PART))
p, T, g = bench.eof_setup()
E = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                  g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                  rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
NSETS = 6
sets = [tuple(ops.dev(a) for a in S.exponential_disc(N, 2002 + k)) for k in range(NSETS)]
outs = [torch.empty((6, N), dtype=torch.float64, device='cuda') for _ in range(NSETS)]
coef = torch.empty((2, g['mmax'] + 1, g['norder']), dtype=torch.float64, device='cuda')
lib = E.lib

def step(k):
    x, y, z, m = sets[k % NSETS]; o = outs[k % NSETS]
    L.check(lib.bfe_eof_prepare(E.h, N, _ptr(x), _ptr(y), _ptr(z), _ptr(m), _stream()))
    L.check(lib.bfe_eof_accumulate_prepared(E.h, _ptr(coef[0]), _ptr(coef[1]), _stream()))
    L.check(lib.bfe_eof_contract(E.h, _ptr(coef[0]), _ptr(coef[1]), 0, g['mmax'], g['norder'], 0, _stream()))
    L.check(lib.bfe_eof_force_prepared(E.h, *[_ptr(o[i]) for i in range(6)], _stream()))

def timed(steps=200):
    for k in range(10): step(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(steps): step(10 + k)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3

step(0); torch.cuda.synchronize()
ref = (coef.clone(), outs[0].clone())
names, vals = [], []
for a in sys.argv[1:]:
    k, v = a.split('='); names.append(k); vals.append([int(t) for t in v.split(',')])
for combo in itertools.product(*vals):
    for k, v in zip(names, combo): ops.set_option(k, v)
    us = min(timed(), timed())
    step(0); torch.cuda.synchronize()
    dc = float((coef - ref[0]).abs().max() / ref[0].abs().max()); do = float((outs[0] - ref[1]).abs().max() / ref[1].abs().max())
    print(json.dumps(dict(zip(names, combo))), 'us/step %.1f' % us, 'particles/s %.3e' % (N / us * 1e6), 'dcoef %.1e dout %.1e' % (dc, do), flush=True)
