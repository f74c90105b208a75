set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest9.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest9.log; tail -30 gpurun_out/r02_pytest9.log | head -40
( time python bench.py --configs none ) > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -c 300 gpurun_out/r02_bench_b.err
