set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest12.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest12.log; tail -25 gpurun_out/r02_pytest12.log | head -40
python profiles/r02_e2e_probe.py > gpurun_out/r02_e2e_probe_b.json 2> gpurun_out/r02_e2e_probe_b.err; tail -12 gpurun_out/r02_e2e_probe_b.json
python - <<'PY'
import sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from exptool_b200 import ops, synthetic as S
from exptool_b200.basis import eof as beof
p, T, g = bench.eof_setup()
geo = (g['XMIN'], g['dX'], g['YMIN'], g['dY'], g['numx'], g['numy'])
tabs = (T['potC'], T['potS'], g['mmax'], g['norder']) + geo + (g['ascale'], g['hscale'], g['cmap'])
E = beof.device_tables(*tabs)
x, y, z, m = S.exponential_disc(1000000, 5)
res = {}
for th in (0, 1, 2, 4, 8):
    ops.set_option('host_threads', th)
    for _ in range(3): E.accumulate_host(x, y, z, m)
    t0 = time.perf_counter()
    for _ in range(10): E.accumulate_host(x, y, z, m)
    res['accumulate_host pageable, host_threads=%d (ms)' % th] = (time.perf_counter() - t0) * 100
print(json.dumps(res, indent=1))
PY
