set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest10.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest10.log; tail -30 gpurun_out/r02_pytest10.log | head -45
( time python bench.py --configs C1 ) > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; tail -c 300 gpurun_out/r02_bench_c.err
python -c "import __graft_entry__ as g; g.smoke()"
