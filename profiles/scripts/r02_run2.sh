set -x
mkdir -p gpurun_out
python profiles/r02_field_probe.py --skip-points --resorts 1,2,3,4 --steps 100 --out gpurun_out/r02_field_probe_b.json > gpurun_out/r02_probe_b.log 2>&1; tail -3 gpurun_out/r02_probe_b.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_orbit_k1.csv python profiles/r02_field_probe.py --norb 1000000 --steps 12 --skip-points --resorts 1 > gpurun_out/ncu3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_points.csv python profiles/r02_field_probe.py --n 1000000 --skip-orbits --chunks 524288 > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:leapfrog_perm_kernel -s 6 -c 1 -o gpurun_out/r02_leapfrog_perm_k1 python profiles/r02_field_probe.py --norb 1000000 --steps 40 --skip-points --resorts 1 > gpurun_out/ncu5.log 2>&1; tail -2 gpurun_out/ncu5.log
ls -la gpurun_out
