set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest1.log 2>&1; tail -5 gpurun_out/r02_pytest1.log
python profiles/r02_field_probe.py --out gpurun_out/r02_field_probe_a.json > gpurun_out/r02_probe_a.log 2>&1; tail -3 gpurun_out/r02_probe_a.log
BFE_LIB=$PWD/exptool_b200/libbfe_minb3.so python profiles/r02_field_probe.py --chunks 524288 --resorts 2,4,8 --out gpurun_out/r02_field_probe_minb3.json > gpurun_out/r02_probe_minb3.log 2>&1; tail -3 gpurun_out/r02_probe_minb3.log
ncu --set full --clock-control none --import-source on -k regex:field_perm_kernel -s 2 -c 2 -o gpurun_out/r02_field_perm python profiles/r02_field_probe.py --n 1000000 --skip-orbits --chunks 524288 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
ncu --set full --clock-control none --import-source on -k regex:leapfrog_perm_kernel -s 3 -c 2 -o gpurun_out/r02_leapfrog_perm python profiles/r02_field_probe.py --norb 500000 --steps 40 --skip-points --resorts 4 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ls -la gpurun_out
