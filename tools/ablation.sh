set -x
B="python bench.py --no-cpu-baseline --no-inference --no-extras --steps 10 --warmup 3"
run() { name=$1; shift; env "$@" $B > gpurun_out/abl_$name.json 2> gpurun_out/abl_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/abl_$name.json").read().strip().splitlines()[-1])
    print("$name", d["value"], d["ms_per_step"], d.get("phases"))
except Exception as e:
    print("$name failed", e)
PY
}
run base X=1
run nop3 HG_OPTIONS=persist_3x3=0
run nowgrad HG_DEBUG_SKIP=hg_conv_wgrad
run nosmallwgrad HG_DEBUG_SKIP=hg_conv_wgrad:@4x4,hg_conv_wgrad:@8x8,hg_conv_wgrad:@16x16
run nobnapply HG_DEBUG_SKIP=hg_bn_apply
run nobnbwd HG_DEBUG_SKIP=hg_bn_bwd_apply
run noups HG_DEBUG_SKIP=hg_upsample2x_bwd,hg_upsample2x_add_fwd
run no3x3 HG_DEBUG_SKIP=hg_conv_fprop_ex:k3,hg_conv_dgrad_bn:k3
run no1x1big HG_DEBUG_SKIP=hg_conv_fprop_ex:k1\ @64x64,hg_conv_dgrad_bn:k1\ @64x64,hg_conv_fprop_ex:k1\ @32x32,hg_conv_dgrad_bn:k1\ @32x32
