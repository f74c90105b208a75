set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest22.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest22.log; tail -5 gpurun_out/r02_pytest22.log
for sp in 0 1; do python profiles/r02_field_probe.py --skip-orbits --chunks 1048576 --opt field_split=$sp --out gpurun_out/r02_field_probe_split$sp.json > gpurun_out/r02_field_probe_split$sp.log 2>&1; done
