set -x
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n2_final2.json 2> gpurun_out/r02_bench_n2_final2.err; tail -c 300 gpurun_out/r02_bench_n2_final2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r02_dist_check_n2_final2.log 2>&1; tail -2 gpurun_out/r02_dist_check_n2_final2.log
