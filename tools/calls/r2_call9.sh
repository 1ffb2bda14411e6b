mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "wgrad_128" 2>&1 | tail -15 > gpurun_out/r2_t9.log
KINDS=wgrad REPS=5 python tools/gpu_top_kernels.py > gpurun_out/r2_top9_kpx64.log 2>&1
HG_OPTIONS=wgrad_kpx=128 KINDS=wgrad REPS=5 python tools/gpu_top_kernels.py > gpurun_out/r2_top9_kpx128.log 2>&1
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-extras"
HG_OPTIONS=wgrad_kpx=128 $B > gpurun_out/r2_bench9_kpx128.json 2> gpurun_out/r2_bench9_kpx128.err
HG_OPTIONS=mid_n_tiles=1 $B > gpurun_out/r2_bench9_mid.json 2> /dev/null
$B > gpurun_out/r2_bench9.json 2> /dev/null
echo; tail -n 4 gpurun_out/r2_t9.log; cat gpurun_out/r2_top9_kpx64.log gpurun_out/r2_top9_kpx128.log
for f in gpurun_out/r2_bench9*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['phases'])
except Exception as e: print('$f ERR', e)"; done
