"""Drop-in for the reference script try_different_stack.py: 3-stack weight-shared hourglass with a different head
per stack -- 2-ch background (CE), 20-ch limb classes (CE), 17-ch keypoint heatmaps (MSE) -- and re-injection by
conv4_k(cat[ll, tmpOut]) (reference try_different_stack.py:24-70,202-329)."""
from ._modules import make_multihead_family
from .targets import gaussian_heatmaps, label_maps

nModules = 2
nFeats = 256
nStack = 3
nKeypoint = 17
nSkeleton = 19
nOutChannels_0 = 2
nOutChannels_1 = nSkeleton + 1
nOutChannels_2 = nKeypoint
epochs = 51
batch_size = 16
keypoints = 17
skeleton = 20
threshold = 0.8

sks = [[15, 13], [13, 11], [16, 14], [14, 12], [11, 12], [5, 11], [6, 12], [5, 6], [5, 7], [6, 8], [7, 9], [8, 10],
       [1, 2], [0, 1], [0, 2], [1, 3], [2, 4], [3, 5], [4, 6]]

ResidualBlock, hourglass, lin, creatModel, _ASPPModule = make_multihead_family(globals())


def render_targets(persons, img_wh, num_persons=None, device="cuda"):
    """(Gauss_map[17,64,64] f32, skeleton map i64, background map i64) of myImageDataset_COCO.__getitem__
    (try_different_stack.py:114-155)."""
    gauss = gaussian_heatmaps(persons, img_wh, J=keypoints, num_persons=num_persons, truncate=True, device=device)
    skel = label_maps(persons, img_wh, sks, num_persons=num_persons, device=device)
    bg = label_maps(persons, img_wh, sks, num_persons=num_persons, line_value=1, device=device)
    return gauss, skel, bg
