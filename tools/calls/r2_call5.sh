mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_t5.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
cp gpurun_out/kernel_table.txt gpurun_out/r2_kernel_table5.txt
export HG_CUDA_GRAPHS=0
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-inference --no-extras > gpurun_out/r2_plain_step.json 2> gpurun_out/r2_plain_step.err && ncu --metrics gpu__time_duration.sum --clock-control none -s 12000 -c 4200 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-inference --no-extras > gpurun_out/r2_ncu_step.log 2>&1
unset HG_CUDA_GRAPHS
echo; tail -n 4 gpurun_out/r2_t5.log; head -c 400 gpurun_out/r2_bench5.json; tail -n 3 gpurun_out/r2_ncu_step.log
