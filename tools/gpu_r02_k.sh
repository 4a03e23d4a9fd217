#!/bin/bash
# Round-2 GPU session K: one default bench line per BASELINE.json configuration (default mode f16f8), reference arm.
mkdir -p gpurun_out
b() { name=$1; shift; timeout -k 5 600 python bench.py "$@" > gpurun_out/r02k_bench_$name.json 2> gpurun_out/r02k_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02k_bench_$name.json | head -n 3; tail -n 1 gpurun_out/r02k_bench_$name.err; }
b dsec_default
b ddd17 --workload ddd17
b uda --workload uda
b bins10 --bins 10 --no-torch-gpu-baseline
b contractA --contract A --no-torch-gpu-baseline
b reference_arm --impl reference --steps 2 --warmup 1
timeout 300 python -m pytest tests/test_gpu_modules.py -q -m gpu -x -k "row_stacked" --timeout 300 -p no:cacheprovider > gpurun_out/r02k_stacked.log 2>&1; echo "stacked tests exit $?"; tail -n 3 gpurun_out/r02k_stacked.log
