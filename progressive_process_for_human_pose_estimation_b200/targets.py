"""Host API of the target-rendering kernels (csrc/targets.cu).

The reference renders targets with numpy / PIL inside DataLoader workers, one image at a time
(try_with_torch.py:107-132, try_different_stack.py:114-155).  Here a whole batch of annotations is rendered by
one kernel launch on the GPU; the arithmetic (float64 Gaussian, Pillow's Bresenham) is restated bit for bit.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L


def _prep(persons, img_wh, num_persons, device):
    persons = torch.as_tensor(np.asarray(persons, dtype=np.float64)) if not torch.is_tensor(persons) else persons
    if persons.dim() == 3:  # [B, J, 3] -> one person per image
        persons = persons.unsqueeze(1)
    if persons.dim() != 4 or persons.shape[-1] != 3:
        raise ValueError("persons must be [B, P, J, 3] (x, y, v)")
    persons = persons.to(device=device, dtype=torch.float64).contiguous()
    B, P, J, _ = persons.shape
    img_wh = torch.as_tensor(np.asarray(img_wh, dtype=np.float64)) if not torch.is_tensor(img_wh) else img_wh
    img_wh = img_wh.to(device=device, dtype=torch.float64).contiguous()
    if tuple(img_wh.shape) != (B, 2):
        raise ValueError("img_wh must be [B, 2] (width, height)")
    if num_persons is not None:
        num_persons = torch.as_tensor(num_persons).to(device=device, dtype=torch.int32).contiguous()
        if tuple(num_persons.shape) != (B,):
            raise ValueError("num_persons must be [B]")
    return persons, img_wh, num_persons, B, P, J


def gaussian_heatmaps(persons, img_wh, J=None, H=64, W=64, num_persons=None, center_mode=0, truncate=True,
                      accumulate=False, pre_scale=1.0, sigma=1.0, amplitude=1.0, device="cuda"):
    """Gaussian keypoint heatmaps [B, J, H, W] float32.

    persons [B,P,J,3] (x, y, v) in image pixels, img_wh [B,2].  Variants:
      truncate=True,  accumulate=False : try_with_torch.py:107-132 (last person wins, quirk Q7)
      truncate=False, pre_scale=100    : try_with_torch_100.py:64-85
      truncate=False                   : only_one_hourgless.py:112-132, read_mscoco.py:46-67
      accumulate=True, center_mode=1   : hourglass_compare.py:713-734 (MPII);  center_mode=0: :286-313 (COCO)
      amplitude=1/(2 pi sigma^2)       : data_argumentation.py:33-52
    """
    persons, img_wh, num_persons, B, P, Jp = _prep(persons, img_wh, num_persons, device)
    if J is not None and J != Jp:
        raise ValueError(f"expected {J} joints, got {Jp}")
    out = torch.empty(B, Jp, H, W, device=persons.device, dtype=torch.float32)
    d = L.HgGaussDesc(B, P, Jp, H, W, center_mode, 1 if truncate else 0, 1 if accumulate else 0, float(pre_scale),
                      float(sigma), float(amplitude))
    with torch.cuda.device(persons.device):
        L.call("hg_render_gauss", C.byref(d), L.ptr(persons), L.ptr(num_persons), L.ptr(img_wh), L.ptr(out),
               L.stream_ptr())
    return out


def label_maps(persons, img_wh, limbs, H=64, W=64, num_persons=None, center_mode=0, draw_points=False,
               draw_lines=True, line_value=0, device="cuda"):
    """Integer label maps [B, H, W] int64 drawn like PIL's ImageDraw.point / ImageDraw.line
    (try_different_stack.py:146-155: skeleton map = limbs with value i+1, background map = limbs with value 1;
    try_skeleton_and_keypoints.py:104-111: keypoint map = points with value k+1).
    draw_points="ellipse" (with center_mode=1): the MPII keypoint map of train.py:668-690, ImageDraw.ellipse on the
    float centre +-0.5; its skeleton map is draw_lines with center_mode=1."""
    persons, img_wh, num_persons, B, P, J = _prep(persons, img_wh, num_persons, device)
    limbs_t = torch.as_tensor(np.asarray(limbs, dtype=np.int32).reshape(-1, 2)).to(persons.device).contiguous()
    if limbs_t.numel() and (int(limbs_t.max()) >= J or int(limbs_t.min()) < 0):
        raise ValueError("limb end point index out of range")
    out = torch.empty(B, H, W, device=persons.device, dtype=torch.int64)
    d = L.HgLabelDesc(B, P, J, limbs_t.shape[0], H, W, center_mode,
                      2 if draw_points == "ellipse" else (1 if draw_points else 0), 1 if draw_lines else 0,
                      int(line_value))
    with torch.cuda.device(persons.device):
        L.call("hg_render_labels", C.byref(d), L.ptr(persons), L.ptr(num_persons), L.ptr(img_wh), L.ptr(limbs_t),
               L.ptr(out), L.stream_ptr())
    return out


def to_tensor_normalize(images_u8, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """transforms.ToTensor() + transforms.Normalize(mean, std) of the reference datasets (try_with_torch.py:310-313)
    for a whole batch on the GPU: uint8 [B,H,W,C] (what PIL hands over after resize) -> fp32 [B,C,H,W], bit-exact
    with torchvision.  The host ships 1 byte per sample instead of 4."""
    import ctypes as C

    if not torch.is_tensor(images_u8) or images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
        raise TypeError("to_tensor_normalize: expected a uint8 tensor [B,H,W,C]")
    if not images_u8.is_cuda:
        raise RuntimeError("to_tensor_normalize: the batch must live on the GPU (there is no CPU fallback)")
    x = images_u8.contiguous()
    B, H, W, Ch = x.shape
    if len(mean) != Ch or len(std) != Ch:
        raise ValueError("to_tensor_normalize: one mean / std per channel")
    out = torch.empty(B, Ch, H, W, device=x.device, dtype=torch.float32)
    m = (C.c_float * Ch)(*[float(v) for v in mean])
    s = (C.c_float * Ch)(*[float(v) for v in std])
    with torch.cuda.device(x.device):
        L.call("hg_image_u8_to_nchw_f32", L.ptr(x), B, H, W, Ch, m, s, L.ptr(out), L.stream_ptr())
    return out


class AnnotationTable:
    """The dataset's annotations, uploaded to HBM once; `batch(indices)` turns a batch of sample indices into the
    [B,P,J,3] keypoint tensor + num_persons + img_wh on the device (one kernel, no host loop) -- the step the reference
    runs per image inside `__getitem__` (`anno.loadAnns(...)`, try_with_torch.py:103-113; `annopoints.point`,
    hourglass_compare.py:691-703).  JSON / .mat parsing and JPEG decoding stay on the host (done once at start-up).

    from_coco(persons_per_image, sizes): persons_per_image[i] = list of flat 3*J `keypoints` lists (ints, as in the
        COCO JSON) of image i, sizes[i] = (width, height).
    from_mpii(points_per_sample, sizes, J=16): points_per_sample[i] = list of (id, x, y, is_visible) records."""

    def __init__(self, table, offsets, sizes, J, mode, device="cuda"):
        self.J, self.mode = int(J), int(mode)
        self.offsets_host = np.asarray(offsets, dtype=np.int32)
        self.table = torch.as_tensor(np.ascontiguousarray(table, dtype=np.float64)).to(device)
        self.offsets = torch.as_tensor(self.offsets_host).to(device)
        self.sizes = torch.as_tensor(np.ascontiguousarray(sizes, dtype=np.float64).reshape(-1, 2)).to(device)
        self.device = self.table.device
        if self.offsets_host[0] != 0 or self.offsets_host[-1] != (self.table.shape[0] if self.table.numel() else 0):
            raise ValueError("AnnotationTable: offsets do not cover the table")
        if len(self.offsets_host) != self.sizes.shape[0] + 1:
            raise ValueError("AnnotationTable: one (width, height) per sample expected")

    @classmethod
    def from_coco(cls, persons_per_image, sizes, J=17, device="cuda"):
        rows, off = [], [0]
        for persons in persons_per_image:
            for kp in persons:
                kp = np.asarray(kp, dtype=np.float64).reshape(-1)
                if kp.size != 3 * J:
                    raise ValueError(f"a COCO person has {3 * J} keypoint numbers, got {kp.size}")
                rows.append(kp)
            off.append(len(rows))
        table = np.stack(rows) if rows else np.zeros([0, 3 * J])
        return cls(table, off, sizes, J, 0, device)

    @classmethod
    def from_mpii(cls, points_per_sample, sizes, J=16, device="cuda"):
        rows, off = [], [0]
        for pts in points_per_sample:
            for rec in pts:
                rows.append(np.asarray(rec, dtype=np.float64).reshape(4))
            off.append(len(rows))
        table = np.stack(rows) if rows else np.zeros([0, 4])
        return cls(table, off, sizes, J, 1, device)

    def __len__(self):
        return len(self.offsets_host) - 1

    def batch(self, indices, max_persons=None):
        """-> (keypoints float64 [B,P,J,3], num_persons int32 [B], img_wh float64 [B,2]), all on the device.
        P = the largest person count of the batch (host arithmetic on the CSR offsets only) unless max_persons is
        given; images with more persons keep their last `max_persons` (the last person wins, quirk Q7)."""
        idx_host = np.asarray(indices, dtype=np.int64).reshape(-1)
        if idx_host.size == 0 or idx_host.min() < 0 or idx_host.max() >= len(self):
            raise IndexError("AnnotationTable.batch: sample index out of range")
        if self.mode == 1:
            P = 1
        else:
            counts = self.offsets_host[idx_host + 1] - self.offsets_host[idx_host]
            P = int(max(1, counts.max())) if max_persons is None else int(max_persons)
        B = idx_host.size
        idx = torch.as_tensor(idx_host).to(self.device)
        kp = torch.empty(B, P, self.J, 3, device=self.device, dtype=torch.float64)
        npers = torch.empty(B, device=self.device, dtype=torch.int32)
        wh = torch.empty(B, 2, device=self.device, dtype=torch.float64)
        d = L.HgAnnotDesc(B, P, self.J, self.mode)
        with torch.cuda.device(self.device):
            L.call("hg_gather_annotations", C.byref(d), L.ptr(self.table), L.ptr(self.offsets), L.ptr(self.sizes),
                   L.ptr(idx), L.ptr(kp), L.ptr(npers), L.ptr(wh), L.stream_ptr())
        return kp, npers, wh
