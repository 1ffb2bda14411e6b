mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_t6.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras --no-inference > gpurun_out/r2_bench6b.json 2> gpurun_out/r2_bench6b.err
echo; tail -n 5 gpurun_out/r2_t6.log; head -c 300 gpurun_out/r2_bench6.json; echo; head -c 300 gpurun_out/r2_bench6b.json
