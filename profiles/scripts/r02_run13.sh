set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 python profiles/sanitize_run.py > gpurun_out/r02_sanitize_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r02_sanitize_$tool.log
done
