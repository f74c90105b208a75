set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest6.log 2>&1; tail -8 gpurun_out/r02_pytest6.log
python profiles/r02_field_probe.py --skip-orbits --chunks 524288,1048576 --out gpurun_out/r02_field_probe_f.json > gpurun_out/r02_probe_f.log 2>&1; tail -3 gpurun_out/r02_probe_f.log
BFE_LIB=$PWD/exptool_b200/libbfe_nopf.so python profiles/r02_field_probe.py --skip-orbits --chunks 524288,1048576 --out gpurun_out/r02_field_probe_f_nopf.json > gpurun_out/r02_probe_f2.log 2>&1; tail -3 gpurun_out/r02_probe_f2.log
ncu --set full --clock-control none --import-source on -k regex:field_rec_kernel -s 2 -c 2 -o gpurun_out/r02_field_rec_f python profiles/r02_field_probe.py --n 1000000 --skip-orbits --chunks 524288 > gpurun_out/ncu6.log 2>&1; tail -2 gpurun_out/ncu6.log
