P=progressive_process_for_human_pose_estimation_b200
cp $P/libhg_sm100a.so /tmp/orig.so
run() { echo "== $1 $2"; cp $P/libhg_$1.so $P/libhg_sm100a.so 2>/dev/null || cp /tmp/orig.so $P/libhg_sm100a.so; BN_ONLY=1 HG_OPTIONS=$2 timeout 120 python tools/gpu_chain_probe.py 2>&1 | tail -3; }
run orig bn_bwd_blocks_per_sm=0
run orig bn_bwd_blocks_per_sm=4
run u3m2 bn_bwd_blocks_per_sm=2
run u2m3 bn_bwd_blocks_per_sm=3
run u2m3 bn_bwd_blocks_per_sm=6
run u2m4 bn_bwd_blocks_per_sm=4
run u8m1 bn_bwd_blocks_per_sm=1
cp /tmp/orig.so $P/libhg_sm100a.so
