set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c5.csv python profiles/prof_c5.py 10000000 > gpurun_out/r02_c5.log 2>&1
tail -2 gpurun_out/r02_c5.log
