mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "persistent" 2>&1 | tail -40 > gpurun_out/r2_t3_p1.log
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_t3.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
cp gpurun_out/kernel_table.txt gpurun_out/r2_kernel_table3.txt
HG_OPTIONS=persist_1x1=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/r2_bench3_nop1.json 2> gpurun_out/r2_bench3_nop1.err
HG_OPTIONS=persist_1x1=0,wgrad_fused_bias=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/r2_bench3_nofb.json 2> gpurun_out/r2_bench3_nofb.err
HG_OPTIONS=persist_1x1=0,bn_apply_u4=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/r2_bench3_nou4.json 2> gpurun_out/r2_bench3_nou4.err
HG_OPTIONS=persist_min_tiles=200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/r2_bench3_p200.json 2> gpurun_out/r2_bench3_p200.err
REPS=5 python tools/gpu_top_kernels.py > gpurun_out/r2_top_events3.log 2>&1
tail -3 gpurun_out/r2_t3_p1.log gpurun_out/r2_t3.log
