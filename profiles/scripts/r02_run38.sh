set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest38.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest38.log
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 --out gpurun_out/r02_field_probe_sub7d.json > gpurun_out/r02_field_probe_sub7d.log 2>&1
python profiles/r02_field_probe.py --skip-orbits --n 16000000 --chunks 4194304 --out gpurun_out/r02_field_probe_sub7e.json > gpurun_out/r02_field_probe_sub7e.log 2>&1
python bench.py --configs C3,C4 --configs-scale 0.125 > gpurun_out/r02_bench_small2.json 2> gpurun_out/r02_bench_small2.err
