set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stable or edges or golden or prepared" > gpurun_out/r02_pytest29.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest29.log
for t in 1024 512; do for s in 3 4; do python bench.py --configs none --streams $s --opt sort_tile_threads=$t > gpurun_out/r02_bench_t${t}_s$s.json 2> gpurun_out/r02_bench_t${t}_s$s.err; done; done
