"""Drop-in for the reference script try_with_torch_100.py: same 4-stack network as try_with_torch.py, Gaussian
targets sharpened x100 with float centres (reference try_with_torch_100.py:64-85) and a PCKh "A" evaluation
each epoch (:283-311,357-367)."""
from ._modules import make_s_family
from .evaluate import make_pckh_a
from .targets import gaussian_heatmaps

nModules = 2
nFeats = 256
nStack = 4
nKeypoint = 17
nOutChannels = nKeypoint
epochs = 1000
batch_size = 16
keypoints = 17

ResidualBlock, hourglass, lin, creatModel = make_s_family(globals())
PCKh = make_pckh_a(globals())


def render_targets(persons, img_wh, device="cuda"):
    """exp(-100 * d^2 / 2) with un-truncated centres, one person per sample (try_with_torch_100.py:64-85)."""
    return gaussian_heatmaps(persons, img_wh, J=keypoints, truncate=False, accumulate=False, pre_scale=100.0,
                             device=device)
