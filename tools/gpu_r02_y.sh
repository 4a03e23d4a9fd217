#!/bin/bash
# Round-2 GPU session Y: final state -- every -m gpu file, smoke, the default bench line (as the driver runs it), step trace.
mkdir -p gpurun_out
bash tools/gpu_ci.sh > gpurun_out/r02y_ci.log 2>&1
echo "ci exit $?"; grep -E "exit [0-9]+|passed|failed|error" gpurun_out/r02y_ci.log | tail -n 16
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/r02y_smoke.log
timeout -k 5 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02y_bench_default.json 2> gpurun_out/r02y_bench_default.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r02y_bench_default.json').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step']), d['e2e'].get('host_wall_ms_per_step'), d['clocks'], 'launches', d['gpu_launches'])
print(d['roofline']['frac'], d['roofline']['mean_launch_ms'], d['roofline']['mma_frac_of_peak'], d.get('cpu_baseline',{}).get('value'), {k:v.get('value') for k,v in d.get('torch_gpu_baseline',{}).items() if isinstance(v,dict)})
PY
timeout 300 python tools/step_trace.py gpurun_out/r02y_step_trace.json > gpurun_out/r02y_step_trace.txt 2>&1; head -n 32 gpurun_out/r02y_step_trace.txt | tail -n 30
