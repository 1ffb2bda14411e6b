# end-of-round-2 re-run of the class ablations with the lean MMA / TMA issue loops (results are wrong with classes
# dropped: timing only)
SC="hg_conv_fprop_ex:@16x16,hg_conv_fprop_ex:@8x8,hg_conv_fprop_ex:@4x4,hg_conv_dgrad_bn:@16x16,hg_conv_dgrad_bn:@8x8,hg_conv_dgrad_bn:@4x4"
SB="hg_bn_apply:M8192,hg_bn_apply:M2048,hg_bn_apply:M512,hg_bn_bwd_apply:M8192,hg_bn_bwd_apply:M2048,hg_bn_bwd_apply:M512"
SW="hg_conv_wgrad:@16x16,hg_conv_wgrad:@8x8,hg_conv_wgrad:@4x4"
bash tools/ablation.sh \
  base X=1 \
  nosmall "HG_DEBUG_SKIP=$SC,$SB,$SW" \
  nobnbwd HG_DEBUG_SKIP=hg_bn_bwd_apply \
  nobnapply HG_DEBUG_SKIP=hg_bn_apply \
  nowgrad HG_DEBUG_SKIP=hg_conv_wgrad \
  no3x3 "HG_DEBUG_SKIP=hg_conv_fprop_ex:k3,hg_conv_dgrad_bn:k3" \
  no1x1 "HG_DEBUG_SKIP=hg_conv_fprop_ex:k1,hg_conv_dgrad_bn:k1"
