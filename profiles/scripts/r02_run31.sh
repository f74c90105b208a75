set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/dist_check.py > gpurun_out/r02_dist_check_n8_final.log 2>&1; tail -3 gpurun_out/r02_dist_check_n8_final.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n8_final.json 2> gpurun_out/r02_bench_n8_final.err; tail -c 300 gpurun_out/r02_bench_n8_final.err
