set -x
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 --configs-scale 0.02 ) > gpurun_out/r02_bench_smoke.json 2> gpurun_out/r02_bench_smoke.err; tail -c 600 gpurun_out/r02_bench_smoke.err
( time python bench.py ) > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 400 gpurun_out/r02_bench_a.err
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest7.log 2>&1; tail -8 gpurun_out/r02_pytest7.log
