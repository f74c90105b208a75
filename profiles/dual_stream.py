"""Throughput of the bench step with S independent particle sets in flight on S streams (one cloned handle per
stream, shared tables) vs one stream:   python profiles/dual_stream.py [nstreams ...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from exptool_b200 import ops, synthetic as S, _lib as L
from exptool_b200.ops import _ptr
import ctypes as C
import bench
N = int(os.environ.get('AB_N', bench.N_PART))
p, T, g = bench.eof_setup()
E0 = ops.EOFTables(T['potC'], T['potS'], g['mmax'], g['norder'], g['XMIN'], g['dX'], g['YMIN'], g['dY'],
                   g['numx'], g['numy'], g['ascale'], g['hscale'], g['cmap'],
                   rforceC=T['rforceC'], zforceC=T['zforceC'], rforceS=T['rforceS'], zforceS=T['zforceS'])
NSETS = 6
sets = [tuple(ops.dev(a) for a in S.exponential_disc(N, 2002 + k)) for k in range(NSETS)]
outs = [torch.empty((6, N), dtype=torch.float64, device='cuda') for _ in range(NSETS)]
lib = E0.lib
def run(ns, steps=240):
    Es = [E0] + [E0.clone() for _ in range(ns - 1)]
    streams = [torch.cuda.Stream() for _ in range(ns)]
    coefs = [torch.empty((2, g['mmax'] + 1, g['norder']), dtype=torch.float64, device='cuda') for _ in range(ns)]
    def step(k):
        i = k % ns
        E, st, coef = Es[i], C.c_void_p(streams[i].cuda_stream), coefs[i]
        x, y, z, m = sets[k % NSETS]; o = outs[k % NSETS]
        L.check(lib.bfe_eof_prepare(E.h, N, _ptr(x), _ptr(y), _ptr(z), _ptr(m), st))
        L.check(lib.bfe_eof_accumulate_prepared(E.h, _ptr(coef[0]), _ptr(coef[1]), st))
        L.check(lib.bfe_eof_contract(E.h, _ptr(coef[0]), _ptr(coef[1]), 0, g['mmax'], g['norder'], 0, st))
        L.check(lib.bfe_eof_force_prepared(E.h, *[_ptr(o[j]) for j in range(6)], st))
    best = 1e30
    for rep in range(3):
        for k in range(12): step(k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        a.record(cur)
        for st in streams: st.wait_stream(cur)
        for k in range(steps): step(12 + k)
        for st in streams: cur.wait_stream(st)
        b.record(cur); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / steps * 1e3)
    return best, outs[(12 + steps - 1) % NSETS].clone(), coefs[(12 + steps - 1) % ns].clone()
ref = None
for ns in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
    us, o, c = run(ns)
    if ref is None: ref = (o, c)
    do = float((o - ref[0]).abs().max() / ref[0].abs().max()); dc = float((c - ref[1]).abs().max() / ref[1].abs().max())
    print(json.dumps({'streams': ns, 'us_per_step': round(us, 1), 'particles_per_s': N / us * 1e6, 'dout': do, 'dcoef': dc}), flush=True)
