BIGC="hg_conv_fprop_ex:@64x64,hg_conv_fprop_ex:@32x32,hg_conv_fprop_ex:@128x128,hg_conv_dgrad_bn:@64x64,hg_conv_dgrad_bn:@32x32,hg_conv_dgrad_bn:@128x128,hg_conv_dgrad:@64x64,hg_conv_dgrad:@32x32,hg_conv_dgrad:@128x128,hg_conv_wgrad:@64x64,hg_conv_wgrad:@32x32,hg_conv_wgrad:@128x128"
BIGB="hg_bn_apply:M131072,hg_bn_apply:M32768,hg_bn_apply:M524288,hg_bn_bwd_apply:M131072,hg_bn_bwd_apply:M32768,hg_bn_bwd_apply:M524288,hg_bn_bwd_reduce,hg_bn_stats"
SP="hg_maxpool2_fwd,hg_maxpool2_bwd,hg_upsample2x_add_fwd,hg_upsample2x_bwd,hg_add,hg_stem_fwd,hg_stem_bwd,hg_nchw_f32_to_nhwc"
bash tools/ablation.sh \
  base X=1 \
  nomaxpool HG_DEBUG_SKIP=hg_maxpool2_fwd,hg_maxpool2_bwd \
  noadd HG_DEBUG_SKIP=hg_add \
  nostem HG_DEBUG_SKIP=hg_stem_fwd,hg_stem_bwd \
  nobnreduce HG_DEBUG_SKIP=hg_bn_bwd_reduce \
  noheads "HG_DEBUG_SKIP=hg_conv_fprop_ex:256->16,hg_conv_fprop_ex:16->256,hg_conv_dgrad:256->16,hg_conv_dgrad:16->256,hg_conv_wgrad:256->16,hg_conv_wgrad:16->256,hg_nchw_f32_to_nhwc" \
  onlysmall HG_DEBUG_SKIP=$BIGC,$BIGB,$SP \
  onlybig "HG_DEBUG_SKIP=hg_conv_fprop_ex:@16x16,hg_conv_fprop_ex:@8x8,hg_conv_fprop_ex:@4x4,hg_conv_dgrad_bn:@16x16,hg_conv_dgrad_bn:@8x8,hg_conv_dgrad_bn:@4x4,hg_conv_wgrad:@16x16,hg_conv_wgrad:@8x8,hg_conv_wgrad:@4x4,hg_bn_apply:M8192,hg_bn_apply:M2048,hg_bn_apply:M512,hg_bn_bwd_apply:M8192,hg_bn_bwd_apply:M2048,hg_bn_bwd_apply:M512"
