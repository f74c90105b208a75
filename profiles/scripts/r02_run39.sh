set -x
mkdir -p gpurun_out
for b in 3 5 6; do python bench.py --configs C4 --opt orbit_key_subbits=$b > gpurun_out/r02_bench_c4_b$b.json 2> gpurun_out/r02_bench_c4_b$b.err; done
for k in 2 4; do python bench.py --configs C4 --opt orbit_resort=$k > gpurun_out/r02_bench_c4_k$k.json 2> gpurun_out/r02_bench_c4_k$k.err; done
