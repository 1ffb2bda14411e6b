set -x
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -m gpu -k "persistent_3x3" 2>&1 | tail -15
export SHAPES=64:128:128:3:0,32:128:128:3:0 KINDS=fprop,dgrad_bn REPS=20
echo "== tile kernel"; HG_OPTIONS=persist_3x3=0 timeout 300 python tools/gpu_top_kernels.py
echo "== persistent untransposed (min_units 256)"; HG_OPTIONS=persist_3x3=1,persist_transposed=0,persist_min_units=256 timeout 300 python tools/gpu_top_kernels.py
echo "== persistent transposed (min_units 256)"; HG_OPTIONS=persist_3x3=1,persist_transposed=1,persist_min_units=256 timeout 300 python tools/gpu_top_kernels.py
