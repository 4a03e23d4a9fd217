#!/bin/bash
# Round-2 GPU session X: ncu capture of the HBM-bound kernels of one steady-state step (DRAM bytes + duration per launch).
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size \
    --clock-control none --kernel-name-base demangled \
    -k 'regex:in_bwd|split_bf16|norm_act|event_prepare|event_stats|pw_conv|task_loss|in_stats|in_finalize|radam|upsample2|wgrad_tc_reduce|pack_weight' \
    -s 400 -c 140 -o gpurun_out/prof_hbm_r02x -f \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-torch-gpu-baseline > gpurun_out/prof_hbm_r02x.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/prof_hbm_r02x.ncu-rep
