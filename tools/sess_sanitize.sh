#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over tools/sanitize_run.py: every kernel of libev2b.so on tiny shapes.
mkdir -p gpurun_out; export PYTHONUNBUFFERED=1
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -1
  grep -E "=========     (at|by|Barrier|Invalid|Race)" gpurun_out/sanitize_$tool.log | head -5
  tail -c 3000 gpurun_out/sanitize_$tool.log > gpurun_out/sanitize_${tool}_tail.log; rm gpurun_out/sanitize_$tool.log
done
