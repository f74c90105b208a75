set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest18.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest18.log; tail -25 gpurun_out/r02_pytest18.log | head -30
for s in 3 2 4; do python bench.py --configs none --streams $s > gpurun_out/r02_bench_f_s$s.json 2> gpurun_out/r02_bench_f_s$s.err; done
