set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest41.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest41.log
BFE_FUZZ_ITERS=60 BFE_FUZZ_SEED=5 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
python bench.py --configs C1,C5 > gpurun_out/r02_bench_c15.json 2> gpurun_out/r02_bench_c15.err
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_run.py > gpurun_out/r02_sanitize4_$tool.log 2>&1; tail -1 gpurun_out/r02_sanitize4_$tool.log; done
