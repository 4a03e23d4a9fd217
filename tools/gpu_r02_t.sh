#!/bin/bash
# Round-2 GPU session T: compute-sanitizer (memcheck, racecheck, initcheck) over every kernel family incl. the round-2 kernels.
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout -k 5 900 compute-sanitizer --tool $tool --error-exitcode 3 python tools/sanitize_smoke.py > gpurun_out/r02t_sanitize_$tool.txt 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok" gpurun_out/r02t_sanitize_$tool.txt | tail -n 12
done
