"""Drop-in for try_skeleton_from_keypoints_merge.py: 4-stack weight-shared hourglass whose 17-channel
keypoint head is extended with 19 limb maps gathered from it (out_skeleton[l] = k[a_l] + k[b_l]); the 36-channel tensor
is what the losses see (MSE on [:, :17], cross-entropy on [:, 17:], reference :406-418) and what conv4 re-injects
(reference :21-70,184-305)."""
from ._modules import make_merge_family
from .targets import gaussian_heatmaps, label_maps

nModules = 2
nFeats = 256
nStack = 4
nKeypoint = 17
nSkeleton = 19
nOutChannels = nKeypoint
epochs = 50
batch_size = 16
keypoints = 17
skeleton = 20

threshold = 0.8

sks = [[15, 13], [13, 11], [16, 14], [14, 12], [11, 12], [5, 11], [6, 12], [5, 6], [5, 7], [6, 8], [7, 9], [8, 10],
       [1, 2], [0, 1], [0, 2], [1, 3], [2, 4], [3, 5], [4, 6]]

ResidualBlock, hourglass, lin, creatModel = make_merge_family(globals())


def render_targets(persons, img_wh, num_persons=None, device="cuda"):
    """(Gaussian keypoint maps float32 [B,17,64,64], skeleton label map int64 [B,64,64]) of
    myImageDataset_COCO.__getitem__ (reference :91-135): Gaussians of the LAST annotated person (quirk Q7), limbs of
    every person drawn with value = limb index (limb 0 draws the background value)."""
    gauss = gaussian_heatmaps(persons, img_wh, num_persons=num_persons, truncate=True, device=device)
    smap = label_maps(persons, img_wh, sks, num_persons=num_persons, line_value=-1, device=device)
    return gauss, smap
