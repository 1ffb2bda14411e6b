set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -m gpu 2>&1 | tail -5
export KINDS=fprop,dgrad_bn,wgrad REPS=20
echo "== defaults (tile kernel 3x3, persistent 1x1)"; timeout 300 python tools/gpu_top_kernels.py
export SHAPES=64:128:128:3:0,32:128:128:3:0 KINDS=fprop,dgrad_bn
echo "== persistent transposed (min_units 256)"; HG_OPTIONS=persist_3x3=1,persist_transposed=1,persist_min_units=256 timeout 300 python tools/gpu_top_kernels.py
echo "== persistent untransposed (min_units 256)"; HG_OPTIONS=persist_3x3=1,persist_transposed=0,persist_min_units=256 timeout 300 python tools/gpu_top_kernels.py
