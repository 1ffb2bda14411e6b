"""Drop-in for train.py, the progressive multi-branch model of the repository (reference train.py:411-601): Q4
residual blocks (:411-447), the executed ASPP modules (`_ASPPModule` :449-461, `ASPP_Block` :465-495), the un-shared
hourglass with stride-2 down-sampling, ASPP bottom and cat skips (:498-540) and the three-stage `creatModel` with
bias-free heads re-injected through torch.cat (:543-601).  Module-global configuration as in the reference.

    import progressive_process_for_human_pose_estimation_b200.train as m
    model = m.creatModel().cuda()
    result = model(images)      # [B,2,64,64] background, [B,16,64,64] limbs, [B,17,64,64] keypoints (autograd-enabled)

The bootstrapped / masked losses of train.py:343-408 keep their class names and `forward(input, target, fraction | mask)`
signatures; each is a few kernel launches (per-pixel NLL / squared error -> radix-select top-k mask -> weighted loss +
gradient) instead of log_softmax + nll_loss + topk + mean and their autograd graph.
"""
import torch.nn as nn

from . import losses as _losses
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

ResidualBlock, _ASPPModule, ASPP_Block, hourglass, creatModel, generateMask = make_train_family(globals())


class PCKh(_PCKhB):
    """train.py:759-791: class-probability input, channel j+1 <-> label value j+1 (evaluator B)."""


class Costomer_CrossEntropyLoss(nn.Module):
    """train.py:343-362: mean of the k = int(H*W*max(fraction, 0.1)) largest per-pixel NLLs of every image."""

    def forward(self, input, target, fraction):
        return _losses.bootstrapped_cross_entropy(input, target, fraction)


class Costomer_CrossEntropyLoss_with_mask(nn.Module):
    """train.py:365-376: mean over all pixels of nll * mask."""

    def forward(self, input, target, mask):
        return _losses.masked_cross_entropy(input, target, mask)


class Costomer_MSELoss_with_mask(nn.Module):
    """train.py:379-391: mean over all elements of (input - target)^2 * mask[B,H,W]."""

    def forward(self, input, target, mask):
        return _losses.masked_mse(input, target, mask)


class Costomer_MSELoss(nn.Module):
    """train.py:394-408: mean of the k = int(H*W*max(fraction, 0.25)) largest squared errors of every image."""

    def forward(self, input, target, fraction):
        return _losses.bootstrapped_mse(input, target, fraction)
