"""Host API of the decode / PCKh kernels (csrc/decode.cu): the reference's evaluators as nn.Modules with the
same call signatures and return types, one kernel launch per batch instead of a Python loop per joint."""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L

_THRESHOLDS = np.arange(0, 0.55, 0.05).astype(np.float32)  # hourglass_compare.py:835, compared in float32


def _as_cuda(t, what):
    if not torch.is_tensor(t):
        raise TypeError(f"{what} must be a torch tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on the GPU: decode / PCKh have no CPU fallback (call .cuda())")
    return t.contiguous()


def decode_argmax(heatmaps):
    """First row-major (y, x) of each map's maximum (hourglass_compare.py:831: torch.nonzero(x >= max)[0]).
    heatmaps [..., H, W] (fp32 / bf16 / fp16) -> (int32 [..., 2] (y, x), float32 [...] max value)."""
    hm = _as_cuda(heatmaps, "heatmaps")
    H, W = hm.shape[-2:]
    n = hm.numel() // (H * W)
    yx = torch.empty(n, 2, device=hm.device, dtype=torch.int32)
    mx = torch.empty(n, device=hm.device, dtype=torch.float32)
    with torch.cuda.device(hm.device):
        L.call("hg_decode_argmax", L.ptr(hm), L.hg_dtype(hm.dtype), n, H, W, L.ptr(yx), L.ptr(mx), L.stream_ptr())
    return yx.view(*hm.shape[:-2], 2), mx.view(hm.shape[:-2])


def pckh_sweep_counts(x, target, rect, chan_offset=0, njoints=None, thresholds=None, absolute=False, logits=False):
    """Integer results of the PCKh threshold sweep, all on the device (no host sync).
    absolute=False: correct when sqrt(d2)/standard < k (evaluators B/C); absolute=True: sqrt(d2) < standard*k
    (evaluator D, calculate_parameters.py:927-929).
    logits=True: x holds the network's raw fp32 class scores and the evaluator sees softmax(x, dim=1) -- the
    `pckh.forward(softmax(result[2]), ...)` call of hourglass_compare.py:1160 without writing the probabilities."""
    x = _as_cuda(x, "x")
    target = _as_cuda(target, "target").to(torch.int64)
    rect = _as_cuda(rect, "rect").to(torch.float32)
    B, Cx, H, W = x.shape
    nj = Cx - chan_offset if njoints is None else njoints
    dev = x.device
    thr = torch.from_numpy(_THRESHOLDS if thresholds is None else np.asarray(thresholds, dtype=np.float32)).to(dev)
    nthr = thr.numel()
    correct = torch.zeros(B, nthr, device=dev, dtype=torch.int32)
    total = torch.zeros(B, nthr, device=dev, dtype=torch.int32)
    predict = torch.zeros(B, nj, 2, device=dev, dtype=torch.int32)
    label = torch.zeros(B, nj, 2, device=dev, dtype=torch.int32)
    found = torch.zeros(B, nj, device=dev, dtype=torch.int32)
    standard = torch.zeros(B, device=dev, dtype=torch.float32)
    if logits:
        if x.dtype != torch.float32:
            raise RuntimeError("pckh_sweep_counts(logits=True): fp32 class scores expected")
        ms = torch.empty(B, H, W, 2, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.call("hg_softmax_stats", L.ptr(x), B, Cx, H, W, L.ptr(ms), L.stream_ptr())
            L.call("hg_pckh_logits", L.ptr(x), L.ptr(ms), 1 if absolute else 0, B, Cx, H, W, L.ptr(target), L.ptr(rect),
                   chan_offset, nj, L.ptr(thr), nthr, L.ptr(correct), L.ptr(total), L.ptr(predict), L.ptr(label),
                   L.ptr(found), L.ptr(standard), L.stream_ptr())
        return dict(correct=correct, total=total, predict=predict, label=label, found=found, standard=standard)
    with torch.cuda.device(dev):
        L.call("hg_pckh_abs" if absolute else "hg_pckh_sweep", L.ptr(x), L.hg_dtype(x.dtype), B, Cx, H, W,
               L.ptr(target), L.ptr(rect), chan_offset, nj, L.ptr(thr), nthr, L.ptr(correct), L.ptr(total), L.ptr(predict), L.ptr(label), L.ptr(found),
               L.ptr(standard), L.stream_ptr())
    return dict(correct=correct, total=total, predict=predict, label=label, found=found, standard=standard)


def _sweep_result(r):
    correct = r["correct"].cpu().numpy().astype(np.float64)
    total = r["total"].cpu().numpy().astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        accuracy = correct / total  # NaN rows when no joint was annotated, like the reference
    pred = r["predict"].cpu().numpy().astype(np.float64)
    lab = r["label"].cpu().numpy().astype(np.float64)
    return accuracy, [p for p in pred], [l for l in lab]


class PCKh_hourglass(nn.Module):
    """PCKh 'C': heatmap channel j <-> label value j+1 (hourglass_compare.py:812-844 `PCKh`,
    performance_compare.py:581-615 `PCKh_hourglass`).  Returns (accuracy[B,11] float64, predicts, labels)."""

    def forward(self, x, target, rect):
        return _sweep_result(pckh_sweep_counts(x, target, rect, 0))


class PCKh_softmax(nn.Module):
    """PCKh 'B': class-probability input, channel j+1 <-> label value j+1, channel 0 = background
    (performance_compare.py:544-578, train.py:759-791).  Returns (accuracy, predicts, labels, stand_dist)."""

    def forward(self, x, target, rect):
        r = pckh_sweep_counts(x, target, rect, 1)
        acc, pred, lab = _sweep_result(r)
        return acc, pred, lab, [s for s in r["standard"].cpu()]


class PCKh_from_logits(nn.Module):
    """PCKh 'B' on raw class scores: `PCKh_from_logits()(result[2], y, rect)` returns what the reference's
    `pckh.forward(nn.functional.softmax(result[2]), y, rect)` returns (hourglass_compare.py:1160,
    performance_compare.py:646-647), the softmax being evaluated inside the decode."""

    def forward(self, x, target, rect):
        r = pckh_sweep_counts(x, target, rect, 1, logits=True)
        acc, pred, lab = _sweep_result(r)
        return acc, pred, lab, [s for s in r["standard"].cpu()]


class PCKh_half_standard(nn.Module):
    """PCKh 'D' (calculate_parameters.py:906-937 `PCKh`): class-probability input (channel j+1 <-> label value j+1),
    ONE threshold: correct when sqrt(d2) < standard * 0.5 with standard = 0.6 * head-box diagonal, float32.
    Returns (accuracy: list of B python floats correct/total, predicts, labels).  Like the reference it raises
    ZeroDivisionError for an image without any annotated joint, and label value C (= x.shape[1]) cannot be
    evaluated (the reference indexes channel C and fails)."""

    def forward(self, x, target, rect):
        nj = x.shape[1] - 1
        r = pckh_sweep_counts(x, target, rect, 1, nj, thresholds=[0.5], absolute=True)
        correct = r["correct"].cpu().numpy()[:, 0]
        total = r["total"].cpu().numpy()[:, 0]
        accuracy = [int(c) / int(t) for c, t in zip(correct, total)]  # ZeroDivisionError when total == 0
        # the reference's predict / label arrays have x.shape[1] rows (row C-1 is never filled)
        pad = np.zeros([x.shape[0], 1, 2])
        pred = np.concatenate([r["predict"].cpu().numpy().astype(np.float64), pad], 1)
        lab = np.concatenate([r["label"].cpu().numpy().astype(np.float64), pad], 1)
        return accuracy, [p for p in pred], [l for l in lab]


def make_pckh_a(g):
    """PCKh 'A' of only_one_hourgless.py:285-313 / try_with_torch_100.py:283-311: loops `batch_size` (module
    global) images and 14 joints, head = channel 13, neck = channel 1; quirk Q8 (x error always 0) is kept so the
    counts are bit-exact.  Returns the python float correct / total."""

    class PCKh(nn.Module):
        def __init__(self):
            super(PCKh, self).__init__()

        def counts(self, x, target):
            x = _as_cuda(x, "x")
            target = _as_cuda(target, "target")
            B = int(g["batch_size"])
            if x.shape[0] < B or target.shape[0] < B:
                raise IndexError(f"PCKh loops batch_size={B} images but got {x.shape[0]}")
            counts = torch.zeros(2, device=x.device, dtype=torch.int32)
            H, W = x.shape[-2:]
            with torch.cuda.device(x.device):
                L.call("hg_pckh_a", L.ptr(x), L.hg_dtype(x.dtype), L.ptr(target), L.hg_dtype(target.dtype), B,
                       x.shape[1], target.shape[1], H, W, 14, 13, 1, L.ptr(counts), L.stream_ptr())
            return counts

        def forward(self, x, target):
            c = self.counts(x, target).cpu()
            return int(c[0]) / int(c[1])

    PCKh.__module__ = g.get("__name__", PCKh.__module__)
    PCKh.__qualname__ = "PCKh"
    return PCKh
