set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest5.log 2>&1; tail -8 gpurun_out/r02_pytest5.log
python profiles/r02_field_probe.py --resorts 1,2,3,4 --steps 100 --chunks 524288,1048576,2097152 --out gpurun_out/r02_field_probe_e.json > gpurun_out/r02_probe_e.log 2>&1; tail -3 gpurun_out/r02_probe_e.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_orbit_e.csv python profiles/r02_field_probe.py --norb 1000000 --steps 14 --skip-points --resorts 2 > gpurun_out/ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_points_e.csv python profiles/r02_field_probe.py --n 1000000 --skip-orbits --chunks 524288 > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leapfrog_perm2_kernel -s 3 -c 1 -o gpurun_out/r02_leapfrog_perm_e python profiles/r02_field_probe.py --norb 1000000 --steps 40 --skip-points --resorts 2 > gpurun_out/ncu5.log 2>&1; tail -2 gpurun_out/ncu5.log
ncu --set full --clock-control none --import-source on -k regex:field_rec2_kernel -s 2 -c 2 -o gpurun_out/r02_field_rec_e python profiles/r02_field_probe.py --n 1000000 --skip-orbits --chunks 524288 > gpurun_out/ncu6.log 2>&1; tail -2 gpurun_out/ncu6.log
