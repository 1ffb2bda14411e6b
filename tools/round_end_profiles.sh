set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 600 gpurun_out/r2f_bench.json; cp gpurun_out/kernel_table.txt gpurun_out/r2f_kernel_table.txt
python tools/gpu_hbm_kernels.py > gpurun_out/r2f_hbm_events.log 2>&1; cp gpurun_out/hbm_kernels_events.json gpurun_out/r2f_hbm_kernels_events.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2f_hbm_ncu.csv python tools/gpu_hbm_kernels.py > gpurun_out/r2f_hbm_ncu.log 2>&1; cp gpurun_out/hbm_kernels_events.json gpurun_out/r2f_hbm_kernels_events_ncu.json
SHAPES=64:128:128:3:0,64:128:256:1:1,64:256:128:1:0 REPS=1 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|conv_wgrad_kernel|conv_persist_kernel" -c 18 -o gpurun_out/r2f_top_kernels python tools/gpu_top_kernels.py > gpurun_out/r2f_top_ncu.log 2>&1
ls -la gpurun_out/r2f_top_kernels.ncu-rep
