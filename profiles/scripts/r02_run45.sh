set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest45.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest45.log; grep -n "^FAILED" gpurun_out/r02_pytest45.log | head
BFE_FUZZ_ITERS=60 BFE_FUZZ_SEED=9 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k key_ordered 2>&1 | tail -2
python profiles/r02_chunk_probe.py 40000000 gpurun_out/r02_chunk_probe_km1.json 7 4194304 > gpurun_out/r02_chunk_probe_km1.log 2>&1; tail -3 gpurun_out/r02_chunk_probe_km1.log
python bench.py --configs C3,C4 > gpurun_out/r02_bench_km1.json 2> gpurun_out/r02_bench_km1.err
python bench.py --configs C3,C4 --opt key_mode=0 > gpurun_out/r02_bench_km0.json 2> gpurun_out/r02_bench_km0.err
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 --out gpurun_out/r02_field_probe_km1.json > gpurun_out/r02_field_probe_km1.log 2>&1
