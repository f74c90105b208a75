set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest17.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest17.log; tail -25 gpurun_out/r02_pytest17.log | head -30
( time python bench.py --configs none ) > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; tail -c 300 gpurun_out/r02_bench_e.err
