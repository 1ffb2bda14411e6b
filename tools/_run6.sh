timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-inference --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['gpu_launches'])"
HG_CUDA_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 11500 -c 4000 --csv --log-file gpurun_out/r2h_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-inference --no-extras > gpurun_out/r2h_launches.log 2>&1
grep -c "at::" gpurun_out/r2h_launches_step.csv; grep "at::" gpurun_out/r2h_launches_step.csv | cut -d, -f5 | cut -c1-120 | sort | uniq -c
