mkdir -p gpurun_out
python -m pytest tests/test_families.py tests/test_gpu_model.py tests/test_gpu_ops.py tests/test_gpu_targets_decode.py -m gpu -q -k "merge or warm or hourglass_module or folded_entry or headline" 2>&1 | tail -80 > gpurun_out/r2_t2.log
for i in 1 2 3; do python -m pytest tests/test_gpu_model.py -m gpu -q -k "hourglass_module" 2>&1 | grep -E "assert|passed|failed|Error" | head -8 >> gpurun_out/r2_t2_hg.log; done
REPS=10 python tools/gpu_hbm_kernels.py > gpurun_out/r2_hbm_events.log 2>&1
cp gpurun_out/hbm_kernels_events.json gpurun_out/r2_hbm_kernels_events.json
REPS=3 python tools/gpu_hbm_kernels.py > /dev/null 2>&1 && REPS=3 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_hbm_ncu.csv python tools/gpu_hbm_kernels.py > gpurun_out/r2_hbm_ncu.log 2>&1
cp gpurun_out/hbm_kernels_events.json gpurun_out/r2_hbm_kernels_events_ncu.json
REPS=5 python tools/gpu_top_kernels.py > gpurun_out/r2_top_events.log 2>&1
export SHAPES=64:128:128:3:0,64:128:256:1:1,64:256:128:1:0 REPS=1
python tools/gpu_top_kernels.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|conv_wgrad_kernel" -c 18 -o gpurun_out/r2_top_kernels python tools/gpu_top_kernels.py > gpurun_out/r2_top_ncu.log 2>&1
ls -la gpurun_out | tail -20
