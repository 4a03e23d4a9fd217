#!/bin/bash
# Round-2 GPU session A: every -m gpu test file, smoke(), and one bench line per BASELINE.json single-GPU config.
mkdir -p gpurun_out
bash tools/gpu_ci.sh > gpurun_out/r02a_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02a_ci.log | tail -n 30
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02a_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/r02a_smoke.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > gpurun_out/r02a_bench_$name.json 2> gpurun_out/r02a_bench_$name.err; echo "bench $name exit $?"; head -c 600 gpurun_out/r02a_bench_$name.json; echo; tail -n 3 gpurun_out/r02a_bench_$name.err; }
b dsec --steps 10 --warmup 3
b ddd17 --workload ddd17 --steps 10 --warmup 3
b contractA --contract A --steps 5 --warmup 3 --no-torch-gpu-baseline
b uda --workload uda --steps 5 --warmup 3
b bins10 --bins 10 --steps 5 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline
