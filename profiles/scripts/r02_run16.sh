set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"eof_" -c 40 -o gpurun_out/r02_step_full -f python profiles/prof_step.py eof 1000000 2 > gpurun_out/r02_step_full.log 2>&1
tail -3 gpurun_out/r02_step_full.log
python profiles/ncu_extract.py gpurun_out/r02_step_full.ncu-rep gpurun_out/r02_ncu_full_step_kernels.csv
