"""TEST INFRASTRUCTURE ONLY -- the two integer passes of Pillow's 8-bit resample (libImaging/Resample.c:
ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc) in numpy, driven by the coefficient tables the product's
host code builds (preprocess.bicubic_coeffs).  tests/test_oracle_targets_pckh.py pins `tables + these passes` to Pillow
itself (Image.resize, the reference's call at try_with_torch.py:99) bit for bit."""
import numpy as np

from progressive_process_for_human_pose_estimation_b200.preprocess import _PRECISION_BITS, bicubic_coeffs


def resize_reference_numpy(img, out_w=256, out_h=256):
    """The two integer passes on the host (numpy), for tests of the coefficient tables: uint8 [h, w, 3] -> [out_h, out_w, 3]."""
    h, w, _ = img.shape
    kx, bx, cx = bicubic_coeffs(w, out_w)
    ky, by, cy = bicubic_coeffs(h, out_h)
    src = img.astype(np.int64)
    tmp = np.empty([h, out_w, 3], dtype=np.uint8)
    for xo in range(out_w):
        x0, n = bx[xo]
        s = (1 << 21) + (src[:, x0:x0 + n, :] * cx[xo, :n].astype(np.int64)[None, :, None]).sum(1)
        tmp[:, xo, :] = np.clip(s >> _PRECISION_BITS, 0, 255)
    t64 = tmp.astype(np.int64)
    out = np.empty([out_h, out_w, 3], dtype=np.uint8)
    for yo in range(out_h):
        y0, n = by[yo]
        s = (1 << 21) + (t64[y0:y0 + n] * cy[yo, :n].astype(np.int64)[:, None, None]).sum(0)
        out[yo] = np.clip(s >> _PRECISION_BITS, 0, 255)
    return out
