set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest25.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest25.log; grep -n "^FAILED\|Error\|assert " gpurun_out/r02_pytest25.log | head -30
python bench.py --configs C3,C4 > gpurun_out/r02_bench_a4.json 2> gpurun_out/r02_bench_a4.err
