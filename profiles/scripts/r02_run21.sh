set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest21.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest21.log
python bench.py --configs none > gpurun_out/r02_bench_i.json 2> gpurun_out/r02_bench_i.err
