"""Drop-in for try_with_aspp_remove_max_pool.py (BASELINE config 4): max-pool replaced by stride-2 residual blocks,
skips merged by cat + 1x1 conv, always-on projection shortcut (quirk Q4), unused ASPP members (quirk Q6), heads of
2 / 20 / 17 channels with cat[inter, ll, tmpOut] re-injection (reference :31-44,153-304)."""
from ._modules import make_nopool_family

nModules = 2
nFeats = 256
nStack = 3
nKeypoint = 17
nSkeleton = 19
nOutChannels_0 = 2
nOutChannels_1 = nSkeleton + 1
nOutChannels_2 = nKeypoint
epochs = 50
batch_size = 32
keypoints = 17
skeleton = 20
threshold = 0.8

sks = [[15, 13], [13, 11], [16, 14], [14, 12], [11, 12], [5, 11], [6, 12], [5, 6], [5, 7], [6, 8], [7, 9], [8, 10],
       [1, 2], [0, 1], [0, 2], [1, 3], [2, 4], [3, 5], [4, 6]]

ResidualBlock, hourglass, lin, creatModel, _ASPPModule = make_nopool_family(globals())
