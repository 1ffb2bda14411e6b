SC="hg_conv_fprop_ex:@16x16,hg_conv_fprop_ex:@8x8,hg_conv_fprop_ex:@4x4,hg_conv_dgrad_bn:@16x16,hg_conv_dgrad_bn:@8x8,hg_conv_dgrad_bn:@4x4"
SB="hg_bn_apply:M8192,hg_bn_apply:M2048,hg_bn_apply:M512,hg_bn_bwd_apply:M8192,hg_bn_bwd_apply:M2048,hg_bn_bwd_apply:M512"
bash tools/ablation.sh \
  swd0 HG_OPTIONS=single_wave_deep=0 \
  wg110 HG_OPTIONS=wgrad_smem_kb=110 \
  swd0_wg110 HG_OPTIONS=single_wave_deep=0,wgrad_smem_kb=110 \
  nosmallconv HG_DEBUG_SKIP=$SC \
  nosmallbn HG_DEBUG_SKIP=$SB \
  no16 "HG_DEBUG_SKIP=hg_conv_fprop_ex:@16x16,hg_conv_dgrad_bn:@16x16,hg_conv_wgrad:@16x16,hg_bn_apply:M8192,hg_bn_bwd_apply:M8192" \
  no8and4 "HG_DEBUG_SKIP=hg_conv_fprop_ex:@8x8,hg_conv_dgrad_bn:@8x8,hg_conv_wgrad:@8x8,hg_bn_apply:M2048,hg_bn_bwd_apply:M2048,hg_conv_fprop_ex:@4x4,hg_conv_dgrad_bn:@4x4,hg_conv_wgrad:@4x4,hg_bn_apply:M512,hg_bn_bwd_apply:M512"
