set -x
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stable or edges or golden or prepared" ) > gpurun_out/r02_pytest20.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest20.log
python bench.py --configs none > gpurun_out/r02_bench_h.json 2> gpurun_out/r02_bench_h.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eof_tile|eof_contract|eof_node" --launch-skip 12 -c 5 -o gpurun_out/r02_sort_full -f python profiles/prof_step.py eof 1000000 4 > gpurun_out/r02_sort_full.log 2>&1
python profiles/ncu_extract.py gpurun_out/r02_sort_full.ncu-rep gpurun_out/r02_ncu_full_sort_kernels.csv
