"""Drop-in for the reference script try_with_torch.py: the canonical 4-stack, weight-shared hourglass with 17
COCO keypoint heatmaps and one MSE loss per stack (reference try_with_torch.py:23-70,179-298).

Configuration is by module globals read at call time, exactly like the reference:

    import progressive_process_for_human_pose_estimation_b200.try_with_torch as m
    m.nStack = 8; m.nOutChannels = 16
    model = m.creatModel().cuda()
    result = model(images)            # list of nStack tensors [B, nOutChannels, 64, 64]
    loss = sum(torch.nn.MSELoss()(r, target) for r in result); loss.backward(); opt.step()
"""
from ._modules import make_s_family
from .targets import gaussian_heatmaps  # noqa: F401  (target rendering of try_with_torch.py:107-132)

nModules = 2
nFeats = 256
nStack = 4
nKeypoint = 17
nSkeleton = 19
nOutChannels = nKeypoint
epochs = 51
batch_size = 16
keypoints = 17
skeleton = 20
threshold = 0.8

sks = [[15, 13], [13, 11], [16, 14], [14, 12], [11, 12], [5, 11], [6, 12], [5, 6], [5, 7], [6, 8], [7, 9], [8, 10],
       [1, 2], [0, 1], [0, 2], [1, 3], [2, 4], [3, 5], [4, 6]]

ResidualBlock, hourglass, lin, creatModel = make_s_family(globals())


def render_targets(persons, img_wh, device="cuda"):
    """Gaussian targets of myImageDataset_COCO.__getitem__ (try_with_torch.py:107-132): integer-truncated centres,
    sigma 1, only the last annotated person survives (quirk Q7).  persons [B,P,17,3], img_wh [B,2]."""
    return gaussian_heatmaps(persons, img_wh, J=keypoints, truncate=True, accumulate=False, device=device)
