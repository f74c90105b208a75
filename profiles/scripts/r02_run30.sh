set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/r02_dist_check_n2_final.log 2>&1; tail -5 gpurun_out/r02_dist_check_n2_final.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err; tail -c 500 gpurun_out/r02_bench_n2_final.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; tail -c 300 gpurun_out/r02_bench_ref_n2.json
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest30.log 2>&1; tail -4 gpurun_out/r02_pytest30.log
