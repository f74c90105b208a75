set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest50.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest50.log; grep -n "^FAILED" gpurun_out/r02_pytest50.log | head
BFE_FUZZ_ITERS=60 BFE_FUZZ_SEED=51 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k key_ordered 2>&1 | tail -2
for o in 1 0; do python bench.py --configs C3 --opt key_smem=$o 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"key_smem $o C3\", d[\"configs\"][\"C3\"][\"ms\"])"; done
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 2>&1 | grep -i "key_order_chunk4194304_us_per_1e6_f32tab0\|only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_points_smem.csv python profiles/prof_field_split.py 0 halo 16777216 > /dev/null 2>&1
