"""Drop-in for the reference script only_one_hourgless.py: single-stack weight-shared hourglass
(reference only_one_hourgless.py:27-60,135-254) with the PCKh "A" evaluator (:285-313) and the LSP float-centre
Gaussian targets (:112-132)."""
from ._modules import make_s_family
from .evaluate import make_pckh_a
from .targets import gaussian_heatmaps

nModules = 2
nFeats = 256
nStack = 1
nOutChannels = 18
epochs = 1000
batch_size = 16
keypoints = 17

ResidualBlock, hourglass, lin, creatModel = make_s_family(globals())
PCKh = make_pckh_a(globals())


def render_targets(persons, img_wh, device="cuda"):
    """LSP Gaussians of myImageDataset.__getitem__ (only_one_hourgless.py:112-132): 14 joints, float centres,
    every joint drawn (the reference does not test visibility: pass v=1)."""
    return gaussian_heatmaps(persons, img_wh, J=14, truncate=False, accumulate=False, device=device)
