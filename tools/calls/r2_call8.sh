mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_t8.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-extras"
$B > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
HG_UNPACK_BARRIER=1 $B > gpurun_out/r2_bench8_barrier.json 2> /dev/null
HG_PACK_SIDE_LANE=0 $B > gpurun_out/r2_bench8_nopacklane.json 2> /dev/null
HG_OPTIONS=wgrad_t1_max_kb=512 $B > gpurun_out/r2_bench8_t1_512.json 2> /dev/null
HG_OPTIONS=bn_bwd_blocks_per_sm=3 $B > gpurun_out/r2_bench8_bnb3.json 2> /dev/null
HG_WGRAD_LANES=2 $B > gpurun_out/r2_bench8_wl2.json 2> /dev/null
HG_WGRAD_LANES=6 $B > gpurun_out/r2_bench8_wl6.json 2> /dev/null
$B > gpurun_out/r2_bench8_again.json 2> /dev/null
echo; tail -n 4 gpurun_out/r2_t8.log
for f in gpurun_out/r2_bench8*.json; do python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d['value'], d['ms_per_step'], d['phases'])
except Exception as e: print('$f ERR', e)"; done
