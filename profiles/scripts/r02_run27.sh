set -x
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final.err
