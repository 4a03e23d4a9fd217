#!/bin/bash
# Round-2 GPU session F (re-entry): full -m gpu suite at HEAD, smoke, default + f16f8 bench lines.
mkdir -p gpurun_out
bash tools/gpu_ci.sh > gpurun_out/r02f_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02f_ci.log | tail -n 20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/r02f_smoke.log
b() { name=$1; shift; timeout -k 5 400 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/r02f_bench_$name.json 2> gpurun_out/r02f_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02f_bench_$name.json; tail -n 2 gpurun_out/r02f_bench_$name.err; }
b f16f8 --mode f16f8 --no-torch-gpu-baseline --no-cpu-baseline --profile-all
b bf16x3 --mode bf16x3 --no-torch-gpu-baseline --no-cpu-baseline --profile-all
