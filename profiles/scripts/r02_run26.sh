set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"eof_" --launch-skip 10 -c 16 -o gpurun_out/r02_step_final -f python profiles/prof_step.py eof 1000000 3 > gpurun_out/r02_step_final.log 2>&1
python profiles/ncu_extract.py gpurun_out/r02_step_final.ncu-rep gpurun_out/r02_ncu_full_step_kernels_final.csv
python profiles/ncu_traffic.py gpurun_out/r02_ncu_full_step_kernels_final.csv profiles/r02_traffic.json > /dev/null; cp profiles/r02_traffic.json gpurun_out/
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest26.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest26.log
( time python bench.py ) > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 200 gpurun_out/r02_bench_final.err
( time python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_bench_ref_final.json 2> gpurun_out/r02_bench_ref_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 6 --warmup 3 --configs none > gpurun_out/r02_bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
