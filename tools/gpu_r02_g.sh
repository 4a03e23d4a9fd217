#!/bin/bash
# Round-2 GPU sessions G and ZF (final code): ncu evidence of the default (f16f8) mode -- one encoder window in full, then the launch list.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_tc' -s 7 -c 7 -o gpurun_out/prof_window_r02zf -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline --windows 3 > gpurun_out/prof_window_r02zf.log 2>&1
echo "window capture exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02zf.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/launches_bench_r02zf.log 2>&1
echo "launch list exit $?"
ls -la gpurun_out | tail -n 8
