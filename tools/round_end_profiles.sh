# Round-end evidence (one B200): GPU tests, the bench line, HBM-kernel table (events + ncu DRAM bytes), ncu --set full of the
# dominant convolution kernels, launch list of one eager step.  PFX names the outputs under gpurun_out/.
PFX=${PFX:-r2g}
set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${PFX}_bench.json 2> gpurun_out/${PFX}_bench.err; tail -c 600 gpurun_out/${PFX}_bench.json; cp gpurun_out/kernel_table.txt gpurun_out/${PFX}_kernel_table.txt
timeout 600 python tools/gpu_hbm_kernels.py > gpurun_out/${PFX}_hbm_events.log 2>&1; cp gpurun_out/hbm_kernels_events.json gpurun_out/${PFX}_hbm_kernels_events.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${PFX}_hbm_ncu.csv python tools/gpu_hbm_kernels.py > gpurun_out/${PFX}_hbm_ncu.log 2>&1; cp gpurun_out/hbm_kernels_events.json gpurun_out/${PFX}_hbm_kernels_events_ncu.json
SHAPES=64:128:128:3:0,64:128:256:1:1,64:256:128:1:0 REPS=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_gemm_kernel|conv_wgrad_kernel|conv_persist_kernel" -c 18 -o gpurun_out/${PFX}_top_kernels python tools/gpu_top_kernels.py > gpurun_out/${PFX}_top_ncu.log 2>&1
ls -la gpurun_out/${PFX}_top_kernels.ncu-rep
HG_CUDA_GRAPHS=0 timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 11500 -c 4000 --csv --log-file gpurun_out/${PFX}_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-inference --no-extras > gpurun_out/${PFX}_launches.log 2>&1
wc -l gpurun_out/${PFX}_launches_step.csv
KINDS=fprop,dgrad_bn,wgrad REPS=20 timeout 300 python tools/gpu_top_kernels.py > gpurun_out/${PFX}_top_kernels_events.txt 2>&1
timeout 300 python tools/gpu_chain_probe.py > gpurun_out/${PFX}_chain_latency.txt 2>&1
