"""Drop-in for the building blocks of train.py that lie on the hot path: the executed ASPP modules
(`_ASPPModule` reference train.py:449-461, `ASPP_Block` :465-495) and the Q4 residual block (:411-447).

train.py's `creatModel` (the progressive multi-branch model, :498-601) and its bootstrapped losses are the next
row N2 of SURVEY 8(f) and are not built; everything here is what that model is assembled from.

    from progressive_process_for_human_pose_estimation_b200.train import ASPP_Block
    aspp = ASPP_Block().cuda()
    y = aspp(x)            # x [B,256,h,w] fp32 NCHW -> [B,256,h,w]; autograd-enabled, same state_dict keys
"""
from ._modules import make_aspp_block, make_q4_block

nModules = 2
nFeats = 256
nStack = 3

ResidualBlock = make_q4_block(globals())
_ASPPModule, ASPP_Block = make_aspp_block(globals())
