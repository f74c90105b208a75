set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "key_subbits or key_ordered or cell_sorted" > gpurun_out/r02_pytest37.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest37.log
python bench.py --configs C3,C4 > gpurun_out/r02_bench_c34c.json 2> gpurun_out/r02_bench_c34c.err
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 --out gpurun_out/r02_field_probe_sub7b.json > gpurun_out/r02_field_probe_sub7b.log 2>&1
python profiles/r02_field_probe.py --skip-orbits --n 16000000 --chunks 4194304 --out gpurun_out/r02_field_probe_sub7c.json > gpurun_out/r02_field_probe_sub7c.log 2>&1
