"""Drop-in for try_skeleton_and_keypoints.py (BASELINE config 3): 4-stack weight-shared hourglass with a 38-channel
head = 18 keypoint classes + 20 limb classes, limb logits mixed in place from the keypoint logits (quirk Q11),
trained with 8 cross-entropy terms on channel slices (reference :21-66,162-301,332-360)."""
from ._modules import make_skeleton_family
from .evaluate import make_pckh_a
from .targets import label_maps

nModules = 2
nFeats = 256
nStack = 4
nKeypoint = 17
nSkeleton = 19
nOutChannels = nKeypoint + nSkeleton + 2
epochs = 50
batch_size = 16
keypoints = 17
skeleton = 20

sks = [[15, 13], [13, 11], [16, 14], [14, 12], [11, 12], [5, 11], [6, 12], [5, 6], [5, 7], [6, 8], [7, 9], [8, 10],
       [1, 2], [0, 1], [0, 2], [1, 3], [2, 4], [3, 5], [4, 6]]

ResidualBlock, hourglass, lin, creatModel = make_skeleton_family(globals())
PCKh = make_pckh_a(globals())


def render_targets(persons, img_wh, num_persons=None, device="cuda"):
    """(keypoint label map, skeleton label map), int64 [B,64,64] (try_skeleton_and_keypoints.py:93-114)."""
    kmap = label_maps(persons, img_wh, sks, num_persons=num_persons, draw_points=True, draw_lines=False, device=device)
    smap = label_maps(persons, img_wh, sks, num_persons=num_persons, device=device)
    return kmap, smap
