set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "key_ordered or cell_sorted or degenerate" ) > gpurun_out/r02_pytest11.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest11.log; tail -20 gpurun_out/r02_pytest11.log | head -30
python profiles/r02_field_probe.py --stage 1 --resorts 1,2,3,4 --steps 100 --chunks 524288,1048576,2097152 --out gpurun_out/r02_field_probe_g1.json > gpurun_out/r02_probe_g1.log 2>&1; tail -3 gpurun_out/r02_probe_g1.log
python profiles/r02_field_probe.py --stage 0 --resorts 2,3 --steps 100 --chunks 1048576 --out gpurun_out/r02_field_probe_g0.json > gpurun_out/r02_probe_g0.log 2>&1; tail -3 gpurun_out/r02_probe_g0.log
ncu --set full --clock-control none --import-source on -k regex:field_stage_kernel -s 2 -c 2 -o gpurun_out/r02_field_stage python profiles/r02_field_probe.py --n 1000000 --skip-orbits --chunks 524288 > gpurun_out/ncu6.log 2>&1; tail -2 gpurun_out/ncu6.log
ncu --set full --clock-control none --import-source on -k regex:leapfrog_stage_kernel -s 3 -c 1 -o gpurun_out/r02_leapfrog_stage python profiles/r02_field_probe.py --norb 1000000 --steps 40 --skip-points --resorts 3 > gpurun_out/ncu5.log 2>&1; tail -2 gpurun_out/ncu5.log
