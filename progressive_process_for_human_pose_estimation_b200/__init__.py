"""B200 (sm_100a) implementation of the stacked-hourglass heatmap-regression hot path of
Xinjie-Qiu/progressive_process_for_human_pose_estimation.

Mirror modules keep the reference's script-level API (module-global configuration, class names, state_dict
keys): try_with_torch, try_with_torch_100, only_one_hourgless, try_different_stack(_without_skeleton), try_with_aspp,
try_with_aspp_remove_max_pool, try_more_layer, try_skeleton_and_keypoints, hourglass_compare, performance_compare,
train.  Kernels live in libhg_sm100a.so (include/hg_sm100a.h); see DESIGN.md.
"""
from ._modules import get_compute_dtype, set_compute_dtype  # noqa: F401
from .evaluate import (PCKh_from_logits, PCKh_half_standard, PCKh_hourglass, PCKh_softmax, decode_argmax,  # noqa: F401
                       pckh_sweep_counts)
from .losses import cross_entropy_losses, mse_losses  # noqa: F401
from .optim import Adam  # noqa: F401
from .preprocess import resize_bicubic  # noqa: F401
from .targets import AnnotationTable, gaussian_heatmaps, label_maps, to_tensor_normalize  # noqa: F401

__all__ = ["set_compute_dtype", "get_compute_dtype", "gaussian_heatmaps", "label_maps", "decode_argmax",
           "pckh_sweep_counts", "PCKh_hourglass", "PCKh_softmax", "PCKh_half_standard", "PCKh_from_logits", "mse_losses", "cross_entropy_losses", "Adam", "to_tensor_normalize", "resize_bicubic", "AnnotationTable"]
