#!/bin/bash
# Round-2 GPU session J: rounds-based pair/classic choice for the N=256 layers -- tc tests + bench.
mkdir -p gpurun_out
bash tools/gpu_ci.sh tests/test_gpu_tc.py tests/test_gpu_modules.py tests/test_gpu_teacher.py > gpurun_out/r02j_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02j_ci.log | tail -n 10
b() { name=$1; shift; timeout -k 5 300 python bench.py --steps 10 --warmup 3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all "$@" > gpurun_out/r02j_bench_$name.json 2> gpurun_out/r02j_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02j_bench_$name.json; tail -n 2 gpurun_out/r02j_bench_$name.err; }
b rounds
ESSB_TC_PAIR=2 b pair_always
python tools/cpu_overhead_probe.py > gpurun_out/r02j_cpu_overhead.txt 2>&1; tail -n 12 gpurun_out/r02j_cpu_overhead.txt
