set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest53.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest53.log; grep -n "^FAILED" gpurun_out/r02_pytest53.log | head
BFE_FUZZ_ITERS=60 BFE_FUZZ_SEED=71 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k key_ordered 2>&1 | tail -2
python bench.py --configs C3,C4 > gpurun_out/r02_bench_slim.json 2> gpurun_out/r02_bench_slim.err
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 2>&1 | grep -A2 "key_order_chunk4194304_us_per_1e6_f32tab0\|only"
python profiles/r02_field_probe.py --skip-orbits --n 16000000 --chunks 4194304 2>&1 | grep -A2 "key_order_chunk4194304_us_per_1e6_f32tab0\|only"
