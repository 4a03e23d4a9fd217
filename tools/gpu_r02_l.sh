#!/bin/bash
# Round-2 GPU session L: CUDA-graph unroll -- tests, then graph vs launch-by-launch benches (DSEC, DDD17), e2e stability check.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_modules.py -q -m gpu -x -k "cuda_graph or row_stacked" --timeout 300 -p no:cacheprovider > gpurun_out/r02l_graph_tests.log 2>&1
rc=$?; echo "graph tests exit $rc"; tail -n 25 gpurun_out/r02l_graph_tests.log
b() { name=$1; shift; timeout -k 5 400 python bench.py --no-torch-gpu-baseline --no-cpu-baseline "$@" > gpurun_out/r02l_bench_$name.json 2> gpurun_out/r02l_bench_$name.err; echo "bench $name exit $?"; python tools/print_bench.py gpurun_out/r02l_bench_$name.json 2>/dev/null | head -n 3; tail -n 2 gpurun_out/r02l_bench_$name.err; }
b graph_default
b nograph_default --no-graph
b graph_s10 --steps 10
b nograph_s10 --steps 10 --no-graph
b ddd17_graph --workload ddd17 --steps 10
b ddd17_nograph --workload ddd17 --steps 10 --no-graph
