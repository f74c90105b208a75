set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest49.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest49.log; grep -n "^FAILED" gpurun_out/r02_pytest49.log | head
BFE_FUZZ_ITERS=60 BFE_FUZZ_SEED=41 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k key_ordered 2>&1 | tail -2
python bench.py --configs C3,C4 > gpurun_out/r02_bench_static.json 2> gpurun_out/r02_bench_static.err
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 --out gpurun_out/r02_field_probe_static.json > gpurun_out/r02_field_probe_static.log 2>&1
python bench.py --configs C3 --configs-scale 0.01 > gpurun_out/r02_bench_static_small.json 2> /dev/null
python bench.py --configs C3 --configs-scale 0.01 --opt field_eval_static=0 > gpurun_out/r02_bench_static_small0.json 2> /dev/null
