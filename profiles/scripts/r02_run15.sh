set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest15.log 2>&1; grep -n "passed\|failed" gpurun_out/r02_pytest15.log; tail -25 gpurun_out/r02_pytest15.log | head -30
( time python bench.py ) > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; tail -c 300 gpurun_out/r02_bench_d.err
( time python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -c 200 gpurun_out/r02_bench_ref.err
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_run.py > gpurun_out/r02_sanitize2_$tool.log 2>&1; tail -2 gpurun_out/r02_sanitize2_$tool.log; done
