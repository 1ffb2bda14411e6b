"""Drop-in for try_more_layer.py: the 4-stack multi-head network whose bottom hourglass level EXECUTES its inline ASPP
(four dilated branches 1/6/12/18 + image-level branch -> cat 1280 -> 1x1 conv, reference try_more_layer.py:249-296) and
whose stacks beyond the third reuse the keypoint head (:339-341)."""
from ._modules import make_multihead_family

nModules = 2
nFeats = 256
nStack = 4
nKeypoint = 17
nSkeleton = 19
nOutChannels_0 = 2
nOutChannels_1 = nSkeleton + 1
nOutChannels_2 = nKeypoint
batch_size = 8
keypoints = 17
skeleton = 20
threshold = 0.8

sks = [[15, 13], [13, 11], [16, 14], [14, 12], [11, 12], [5, 11], [6, 12], [5, 6], [5, 7], [6, 8], [7, 9], [8, 10],
       [1, 2], [0, 1], [0, 2], [1, 3], [2, 4], [3, 5], [4, 6]]

ResidualBlock, hourglass, lin, creatModel, _ASPPModule = make_multihead_family(globals(), aspp_members=True,
                                                                               aspp_executed=True)
