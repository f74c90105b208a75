set -x
mkdir -p gpurun_out
for i in 1 2 3; do python bench.py --configs C5 > gpurun_out/r02_bench_wave$i.json 2> gpurun_out/r02_bench_wave$i.err; done
