set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest44.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest44.log
( time python bench.py ) > gpurun_out/r02_bench_final3.json 2> gpurun_out/r02_bench_final3.err; tail -c 200 gpurun_out/r02_bench_final3.err
( time python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_bench_ref_final3.json 2> gpurun_out/r02_bench_ref_final3.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke3.log 2>&1; tail -1 gpurun_out/r02_smoke3.log
