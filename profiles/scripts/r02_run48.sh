set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest48.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest48.log; grep -n "^FAILED" gpurun_out/r02_pytest48.log | head
BFE_FUZZ_ITERS=60 BFE_FUZZ_SEED=31 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k key_ordered 2>&1 | tail -2
python bench.py --configs C3 > gpurun_out/r02_bench_agg1.json 2> gpurun_out/r02_bench_agg1.err
python bench.py --configs C3 --opt key_agg=0 > gpurun_out/r02_bench_agg0.json 2> gpurun_out/r02_bench_agg0.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_points_agg.csv python profiles/prof_field_split.py 0 halo 16777216 > /dev/null 2>&1
