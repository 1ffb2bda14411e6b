mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_t7.log
python -m pytest tests -m gpu -q 2>&1 | grep -E "^(FAILED|ERROR)" | head -20 > gpurun_out/r2_t7_failed.log
echo; tail -n 5 gpurun_out/r2_t7.log; cat gpurun_out/r2_t7_failed.log
