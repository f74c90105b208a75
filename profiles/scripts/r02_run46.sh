set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest46.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest46.log; grep -n "^FAILED" gpurun_out/r02_pytest46.log | head
BFE_FUZZ_ITERS=60 BFE_FUZZ_SEED=11 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k key_ordered 2>&1 | tail -2
python bench.py --configs C3,C4 > gpurun_out/r02_bench_km2.json 2> gpurun_out/r02_bench_km2.err
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 --out gpurun_out/r02_field_probe_km2.json > gpurun_out/r02_field_probe_km2.log 2>&1
python profiles/r02_field_probe.py --skip-orbits --n 16000000 --chunks 4194304 --out gpurun_out/r02_field_probe_km2b.json > gpurun_out/r02_field_probe_km2b.log 2>&1
