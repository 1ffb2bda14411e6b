mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "persistent or upsample" 2>&1 | tail -40 > gpurun_out/r2_t4_p1.log
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_t4.log
REPS=5 python tools/gpu_top_kernels.py > gpurun_out/r2_top_events4.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
cp gpurun_out/kernel_table.txt gpurun_out/r2_kernel_table4.txt
HG_OPTIONS=persist_1x1=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/r2_bench4_nop1.json 2> gpurun_out/r2_bench4_nop1.err
REPS=10 python tools/gpu_hbm_kernels.py > gpurun_out/r2_hbm_events4.log 2>&1
echo; tail -n 3 gpurun_out/r2_t4_p1.log gpurun_out/r2_t4.log
