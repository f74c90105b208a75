set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest36.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest36.log; grep -n "^FAILED" gpurun_out/r02_pytest36.log | head
python bench.py --configs C3,C4 > gpurun_out/r02_bench_c34b.json 2> gpurun_out/r02_bench_c34b.err
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 --out gpurun_out/r02_field_probe_sub7.json > gpurun_out/r02_field_probe_sub7.log 2>&1
