set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest42.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest42.log
python bench.py --configs C5 > gpurun_out/r02_bench_wave.json 2> gpurun_out/r02_bench_wave.err
