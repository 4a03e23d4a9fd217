#!/bin/bash
# Round-2 GPU session R: merged sub-pixel phases of the transposed convolutions -- tests, then contract A / B benches with and without.
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py tests/test_golden_tc.py > gpurun_out/r02r_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02r_ci.log | tail -n 8
grep -E "^(FAILED|ERROR)" gpurun_out/test_gpu_modules.log | head
b() { name=$1; shift; timeout -k 5 400 python bench.py --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all "$@" > gpurun_out/r02r_bench_$name.json 2> gpurun_out/r02r_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02r_bench_$name.json 2>/dev/null | head -n 6; tail -n 1 gpurun_out/r02r_bench_$name.err; }
b A_merged --contract A
ESS_B200_MERGE_PHASES=0 b A_phases --contract A
b B_merged
ESS_B200_MERGE_PHASES=0 b B_phases
