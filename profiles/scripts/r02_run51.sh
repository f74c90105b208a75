set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest51.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest51.log; grep -n "^FAILED" gpurun_out/r02_pytest51.log | head
BFE_FUZZ_ITERS=40 BFE_FUZZ_SEED=61 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
python bench.py --configs C3,C4 > gpurun_out/r02_bench_fin4.json 2> gpurun_out/r02_bench_fin4.err
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 2>&1 | grep -A2 "key_order_chunk4194304_us_per_1e6_f32tab0\|only"
