"""Drop-in for hourglass_compare.py: the un-shared 4-stage 'stacked hourglass' baseline on MPII-16 (nearest
up-sampling, always-on projection shortcut, bias-free heads), its MPII Gaussian targets and the PCKh threshold
sweep (reference hourglass_compare.py:37-51,405-638,713-734,812-844)."""
import torch.nn as nn

from ._modules import make_u_family
from .evaluate import PCKh_hourglass
from .targets import gaussian_heatmaps

nModules = 2
nFeats = 256
nStack = 3
nKeypoint = 16
nSkeleton = 19
nOutChannels_0 = 2
nOutChannels_1 = 16
nOutChannels_2 = 17
batch_size = 30
keypoints = 16
skeleton = 20
inputsize = 256
threshold = 1

ResidualBlock, hourglass, creatModel = make_u_family(globals())


class PCKh(PCKh_hourglass):
    """hourglass_compare.py:812-844: forward(x, target, rect) -> (accuracy[B,11], predicts, labels)."""


def render_targets(points, img_wh, device="cuda"):
    """MPII Gaussians accumulated with `+=`, float centres x * 256 / w / 4 (hourglass_compare.py:713-734).
    points [B,16,3] = (x, y, visible)."""
    return gaussian_heatmaps(points, img_wh, J=16, truncate=False, accumulate=True, center_mode=1, device=device)
