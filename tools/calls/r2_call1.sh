mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -q -k "stride2 or fails_loudly" 2>&1 | tail -30 > gpurun_out/r2_t_stride.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_ops.py::test_conv_stride2_fprop_dgrad_wgrad 2>&1 | tail -60 > gpurun_out/r2_t_all.log
python tools/gpu_hbm_kernels.py > gpurun_out/r2_hbm_events.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
cp gpurun_out/kernel_table.txt gpurun_out/r2_kernel_table1.txt
for f in 3 4 5; do HG_FOLD_BN=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference > gpurun_out/r2_bench1_fold$f.json 2> gpurun_out/r2_bench1_fold$f.err; done
python bench.py --impl torch-eager --steps 5 > gpurun_out/r2_eager.json 2> gpurun_out/r2_eager.err
tail -5 gpurun_out/r2_t_stride.log gpurun_out/r2_t_all.log
