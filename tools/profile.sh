#!/bin/bash
# ncu evidence for the round (run under gpurun, 1 GPU).  Outputs land in gpurun_out/; summaries are
# copied into profiles/ by tools/summarize_profiles.py on the build box.
mkdir -p gpurun_out
R=${ROUND:-r01}
# (1) every launch of one bench run with its device time (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/launches_bench_$R.log 2>&1
echo "launch list exit $?"
# (2) full capture of the dominant kernel (fused ConvLSTM cell on tcgen05): the three pyramid levels of one
#     window after warm-up (conv_tc_kernel<1> = LSTM epilogue instantiation)
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:conv_tc_kernel<.int.1>' -s 9 -c 3 \
    -o gpurun_out/prof_lstm_$R -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --windows 4 \
    > gpurun_out/prof_lstm_$R.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/
