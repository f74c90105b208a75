set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest33.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest33.log; grep -n "^FAILED\|Error\|assert " gpurun_out/r02_pytest33.log | head -20
for st in 1 0; do python bench.py --configs C1,C5 --opt sort_stable=$st > gpurun_out/r02_bench_sl_st$st.json 2> gpurun_out/r02_bench_sl_st$st.err; done
