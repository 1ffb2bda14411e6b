"""Drop-in for train.py, the progressive multi-branch model of the repository (reference train.py:411-601): Q4
residual blocks (:411-447), the executed ASPP modules (`_ASPPModule` :449-461, `ASPP_Block` :465-495), the un-shared
hourglass with stride-2 down-sampling, ASPP bottom and cat skips (:498-540) and the three-stage `creatModel` with
bias-free heads re-injected through torch.cat (:543-601).  Module-global configuration as in the reference.

    import progressive_process_for_human_pose_estimation_b200.train as m
    model = m.creatModel().cuda()
    result = model(images)      # [B,2,64,64] background, [B,16,64,64] limbs, [B,17,64,64] keypoints (autograd-enabled)

The bootstrapped / masked losses of train.py:343-391 are NOT built (they remain stock PyTorch on the returned maps).
"""
from ._modules import make_train_family
from .evaluate import PCKh_softmax as _PCKhB

nModules = 2
nFeats = 256
nStack = 3
nKeypoint_COCO = 17
nSkeleton_COCO = 19
nKeypoint_MPII = 16
nSkeleton_MPII = 15
nOutChannels_0 = 2
nOutChannels_1 = nSkeleton_MPII + 1
nOutChannels_2 = nKeypoint_MPII + 1
batch_size = 48
keypoints = 17
skeleton = 20
inputsize = 256
threshold = 1

ResidualBlock, _ASPPModule, ASPP_Block, hourglass, creatModel = make_train_family(globals())


class PCKh(_PCKhB):
    """train.py:759-791: class-probability input, channel j+1 <-> label value j+1 (evaluator B)."""
