set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest34.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest34.log
( time python bench.py ) > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; tail -c 200 gpurun_out/r02_bench_final2.err
for tool in memcheck racecheck synccheck; do timeout 1200 compute-sanitizer --tool $tool python profiles/sanitize_run.py > gpurun_out/r02_sanitize3_$tool.log 2>&1; tail -2 gpurun_out/r02_sanitize3_$tool.log; done
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke2.log 2>&1; tail -2 gpurun_out/r02_smoke2.log
