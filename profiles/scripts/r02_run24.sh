set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest24.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest24.log; grep -n "^FAILED\|Error\|assert " gpurun_out/r02_pytest24.log | head -30
python profiles/r02_field_probe.py --chunks 1048576 --resorts 3,4 --out gpurun_out/r02_field_probe_a4.json > gpurun_out/r02_field_probe_a4.log 2>&1
