"""Input pipeline on the GPU: `image.resize([256, 256])` (PIL default filter BICUBIC) + `ToTensor()` +
`Normalize(0.5, 0.5)` of the reference datasets (try_with_torch.py:96-101,310-313) for a batch of variable-size RGB
images.  The host only computes Pillow's coefficient tables (they depend on the image size alone and are cached); all
pixel work is two integer kernels (csrc/resize.cu), bit-exact with Pillow's Resample.c.  JPEG decoding stays with the
caller (PIL / nvJPEG): this module takes decoded uint8 HWC pixels that already live on the GPU.
"""
import ctypes as C
import functools
import math

import numpy as np
import torch

from . import _lib as L

_PRECISION_BITS = 32 - 8 - 2   # Resample.c


def _bicubic(x):
    """bicubic_filter of Pillow's Resample.c (a = -0.5), same operation order, float64."""
    x = np.abs(x)
    a = -0.5
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


@functools.lru_cache(maxsize=4096)
def bicubic_coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc of Pillow's Resample.c for one axis and the full-image box:
    (ksize, bounds int32 [out, 2] = (first input index, count), coefficients int32 [out, ksize], 22-bit fixed point)."""
    in0, in1 = 0.0, float(np.float32(in_size))
    scale = (in1 - in0) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = in0 + (xx + 0.5) * scale
    ss = 1.0 / filterscale
    xmin = np.trunc(center - support + 0.5).astype(np.int64)      # (int) casts truncate toward zero
    xmin = np.maximum(xmin, 0)
    xmax = np.trunc(center + support + 0.5).astype(np.int64)
    xmax = np.minimum(xmax, in_size) - xmin
    k = np.zeros([out_size, ksize], dtype=np.float64)
    ww = np.zeros(out_size, dtype=np.float64)
    for x in range(ksize):                                         # sequential accumulation, like the C loop
        valid = x < xmax
        w = _bicubic((((x + xmin).astype(np.float64) - center) + 0.5) * ss)
        w = np.where(valid, w, 0.0)
        k[:, x] = w
        ww = np.where(valid, ww + w, ww)
    nz = ww != 0.0
    k[nz] = k[nz] / ww[nz, None]
    scaled = k * float(1 << _PRECISION_BITS)
    fixed = np.where(k < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int32)
    fixed[np.arange(ksize)[None, :] >= xmax[:, None]] = 0
    bounds = np.stack([xmin, xmax], 1).astype(np.int32)
    return ksize, bounds, fixed


def resize_bicubic(images, size=(256, 256), normalize=None):
    """`[Image.resize(size) for image in batch]` on the GPU.  images: list of uint8 CUDA tensors [h_i, w_i, 3] (decoded
    RGB pixels).  Returns uint8 [B, H, W, 3]; with normalize=(mean, std) returns instead the fp32 [B, 3, H, W] tensor
    `Normalize(mean, std)(ToTensor()(resized))` the reference feeds its models (try_with_torch.py:310-313)."""
    if len(images) == 0:
        raise ValueError("resize_bicubic: empty batch")
    out_w, out_h = int(size[0]), int(size[1])
    dev = images[0].device
    coefs, bounds, descs, keep = [], [], [], []
    coff = boff = toff = 0
    max_h = 0
    for im in images:
        if not torch.is_tensor(im) or im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
            raise TypeError("resize_bicubic: uint8 tensors [h, w, 3] expected")
        if not im.is_cuda:
            raise RuntimeError("resize_bicubic: the decoded pixels must live on the GPU (there is no CPU fallback)")
        im = im.contiguous()
        keep.append(im)
        h, w = int(im.shape[0]), int(im.shape[1])
        kx, bx, cx = bicubic_coeffs(w, out_w)
        ky, by, cy = bicubic_coeffs(h, out_h)
        descs.append(L.HgResizeImage(im.data_ptr(), w, h, coff, coff + cx.size, boff, boff + bx.size, kx, ky, toff))
        coefs += [cx.reshape(-1), cy.reshape(-1)]
        bounds += [bx.reshape(-1), by.reshape(-1)]
        coff += cx.size + cy.size
        boff += bx.size + by.size
        toff += h * out_w * 3
        max_h = max(max_h, h)
    coef_d = torch.from_numpy(np.concatenate(coefs)).to(dev)
    bounds_d = torch.from_numpy(np.concatenate(bounds)).to(dev)
    arr = (L.HgResizeImage * len(descs))(*descs)
    desc_d = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    tmp = torch.empty(toff, device=dev, dtype=torch.uint8)
    B = len(images)
    if normalize is None:
        out = torch.empty(B, out_h, out_w, 3, device=dev, dtype=torch.uint8)
        o8, on, m, s = L.ptr(out), None, None, None
    else:
        out = torch.empty(B, 3, out_h, out_w, device=dev, dtype=torch.float32)
        m = (C.c_float * 3)(*[float(v) for v in normalize[0]])
        s = (C.c_float * 3)(*[float(v) for v in normalize[1]])
        o8, on = None, L.ptr(out)
    with torch.cuda.device(dev):
        L.call("hg_resize_bicubic_u8", L.ptr(desc_d), B, max_h, out_w, out_h, L.ptr(coef_d), L.ptr(bounds_d), L.ptr(tmp),
               o8, on, m, s, L.stream_ptr())
    return out
