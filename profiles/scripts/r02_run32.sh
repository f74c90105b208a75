set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "key_ordered or cell_sorted" > gpurun_out/r02_pytest32.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest32.log
python bench.py --configs C3,C4 --configs-scale 0.125 > gpurun_out/r02_bench_small.json 2> gpurun_out/r02_bench_small.err
python bench.py --configs C3,C4 > gpurun_out/r02_bench_c34.json 2> gpurun_out/r02_bench_c34.err
