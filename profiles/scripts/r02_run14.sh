set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest14.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest14.log; tail -25 gpurun_out/r02_pytest14.log | head -40
for st in 1 0 1 0; do python bench.py --configs none --no-cpu --sort-stable $st > gpurun_out/r02_bench_stable$st.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r02_bench_stable$st.json').read().strip().splitlines()[-1])
print('sort_stable=$st', 'ms/step', round(d['ms_per_step'],4), 'single', round(d['single_stream']['ms_per_step'],4), {k: round(v*1e3,1) for k,v in d['roofline']['kernel_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],3))
PY
done
