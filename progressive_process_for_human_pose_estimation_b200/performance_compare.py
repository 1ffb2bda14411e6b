"""Drop-in for the evaluation script performance_compare.py: the hourglass baseline network
(`creatModel_hourglass`, reference :335-427) and both PCKh evaluators -- `PCKh` on class-probability maps with a
background channel (:544-578) and `PCKh_hourglass` on heatmaps (:581-615)."""
from ._modules import make_u_family
from .evaluate import PCKh_hourglass as _PCKhC
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
batch_size = 30
keypoints = 17
skeleton = 20
inputsize = 256
threshold = 1

ResidualBlock, hourglass, creatModel_hourglass = make_u_family(globals())


class PCKh(_PCKhB):
    """performance_compare.py:544-578: returns (accuracy, predicts, labels, stand_dist)."""


class PCKh_hourglass(_PCKhC):
    """performance_compare.py:581-615: returns (accuracy, predicts, labels, stand_dist) -- the reference appends the
    list to itself (`stand_dist.append(stand_dist)`, :614); here the per-image `standard` values are returned."""

    def forward(self, x, target, rect):
        from .evaluate import _sweep_result, pckh_sweep_counts

        r = pckh_sweep_counts(x, target, rect, 0)
        acc, pred, lab = _sweep_result(r)
        return acc, pred, lab, [s for s in r["standard"].cpu()]
