"""Drop-in for try_different_stack_without_skeleton.py: the 2-stack variant (2-ch background CE + 17-ch keypoint
MSE) of try_different_stack.py (reference try_different_stack_without_skeleton.py:24-36,282-320)."""
from ._modules import make_multihead_family

nModules = 2
nFeats = 256
nStack = 2
nKeypoint = 17
nSkeleton = 19
nOutChannels_0 = 2
nOutChannels_1 = nKeypoint
batch_size = 16
keypoints = 17
skeleton = 20
threshold = 0.8

sks = [[15, 13], [13, 11], [16, 14], [14, 12], [11, 12], [5, 11], [6, 12], [5, 6], [5, 7], [6, 8], [7, 9], [8, 10],
       [1, 2], [0, 1], [0, 2], [1, 3], [2, 4], [3, 5], [4, 6]]

ResidualBlock, hourglass, lin, creatModel, _ASPPModule = make_multihead_family(globals(), num_heads=2)
