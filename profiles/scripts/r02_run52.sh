set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q -k "key_ordered or cell_sorted or key_subbits" 2>&1 | tail -2
for cfg in "field_support_slim=1" "field_support_slim=0"; do python bench.py --configs C3,C4 --opt $cfg 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(\"$cfg C3\", d[\"configs\"][\"C3\"][\"ms\"], \"C4\", d[\"configs\"][\"C4\"][\"ms\"], d[\"configs\"][\"C3\"][\"parity\"][\"key_ordered_equals_caller_order_bits\"])"; done
python profiles/r02_field_probe.py --skip-orbits --chunks 4194304 2>&1 | grep -A2 "key_order_chunk4194304_us_per_1e6_f32tab0\|only"
