"""Fused intermediate-supervision losses (csrc/loss.cu): per-stack MSE and per-pixel class cross-entropy heads.

The reference sums one `nn.MSELoss` per stack (try_with_torch.py:305-308,333-341); those stock modules keep working
on the heatmaps the drop-in models return.  `mse_losses(result, target)` is the optional fast path: one kernel reads
the target once and every stack's prediction once, produces all per-stack losses and, in the same sweep, the
gradient tensors the backward pass needs.
"""
import ctypes as C

import torch

from . import _lib as L


class _MseMulti(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, *preds):
        S = len(preds)
        if S < 1 or S > 8:
            raise RuntimeError("mse_losses: 1..8 stacks supported")
        for p in preds:
            if not p.is_cuda or p.dtype != torch.float32 or p.shape != target.shape:
                raise RuntimeError("mse_losses: predictions must be fp32 CUDA tensors of the target's shape "
                                   "(there is no CPU fallback)")
        preds = [p.contiguous() for p in preds]
        target = target.contiguous().to(torch.float32)
        need = [ctx.needs_input_grad[i + 1] for i in range(S)]
        grads = [torch.empty_like(p) if n else None for p, n in zip(preds, need)]
        loss = torch.empty(S, device=target.device, dtype=torch.float32)
        d = L.HgMseDesc(target.numel(), S, 1.0)
        parr = (C.c_void_p * S)(*[p.data_ptr() for p in preds])
        garr = (C.c_void_p * S)(*[g.data_ptr() if g is not None else None for g in grads])
        with torch.cuda.device(target.device):
            L.zero_(loss)
            L.call("hg_mse_multi", C.byref(d), parr, L.ptr(target), garr, L.ptr(loss), L.stream_ptr())
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, gloss):
        grads = ctx.grads
        ctx.grads = None
        live = [g for g in grads if g is not None]
        if live:
            # d(losses[s]) scales the gradient hg_mse_multi already wrote for stack s: one launch, in place
            numel = live[0].numel()
            if numel % 4 == 0 and gloss.is_cuda:
                gl = gloss.detach().to(torch.float32).contiguous()
                arr = (C.c_void_p * len(grads))(*[g.data_ptr() if g is not None else None for g in grads])
                with torch.cuda.device(gl.device):
                    L.call("hg_scale_multi", numel, len(grads), arr, L.ptr(gl), L.stream_ptr())
            else:
                grads = [None if g is None else g * gloss[s] for s, g in enumerate(grads)]
        return (None, *grads)


def mse_losses(result, target):
    """Per-stack MSE losses as one tensor [nStack]; `mse_losses(result, y).sum().backward()` is the training
    objective of try_with_torch.py:333-342."""
    return _MseMulti.apply(target, *result)


def _plane_view(t):
    """True when `t` [B,C,H,W] can be read in place: channel planes contiguous and H*W apart (a channel slice of a
    contiguous NCHW tensor qualifies)."""
    B, C, H, W = t.shape
    sb, sc, sh, sw = t.stride()
    return sw == 1 and sh == W and sc == H * W and sb >= C * H * W


class _CeMulti(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, *logits):
        # spec: (terms, ignore_index, check_labels); terms[t] = (index into logits, c0, c1, target)
        terms, ignore_index, check = spec
        T = len(terms)
        if T < 1 or T > L.HG_CE_MAX_TERMS:
            raise RuntimeError(f"cross_entropy_losses: 1..{L.HG_CE_MAX_TERMS} terms supported")
        B, _, H, W = logits[0].shape
        HW = H * W
        dev = logits[0].device
        srcs, grads = [], []
        for i, x in enumerate(logits):
            if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or x.shape[0] != B or x.shape[2:] != (H, W):
                raise RuntimeError("cross_entropy_losses: logits must be fp32 CUDA tensors [B,C,H,W] of one batch / "
                                   "map size (there is no CPU fallback)")
            srcs.append(x if _plane_view(x) else x.contiguous())
            if ctx.needs_input_grad[i + 1]:
                covered = sorted((c0, c1) for k, c0, c1, _ in terms if k == i)
                full = covered[0][0] == 0 and covered[-1][1] == x.shape[1] and all(
                    a[1] == b[0] for a, b in zip(covered, covered[1:]))
                grads.append((torch.empty if full else torch.zeros)(x.shape, device=dev, dtype=torch.float32))
            else:
                grads.append(None)
        arr = (L.HgCeTerm * T)()
        keep = []
        for t, (k, c0, c1, target) in enumerate(terms):
            x, g = srcs[k], grads[k]
            if not (0 <= c0 < c1 <= x.shape[1]):
                raise RuntimeError(f"cross_entropy_losses: channel slice [{c0}, {c1}) outside {x.shape[1]} channels")
            if not target.is_cuda or tuple(target.shape) != (B, H, W):
                raise RuntimeError("cross_entropy_losses: labels must be a CUDA tensor [B,H,W]")
            tg = target.contiguous().to(torch.int64)
            keep.append(tg)
            arr[t].logits = x.data_ptr() + 4 * c0 * HW
            arr[t].dlogits = (g.data_ptr() + 4 * c0 * HW) if g is not None else None
            arr[t].target = tg.data_ptr()
            arr[t].logits_bstride = x.stride(0)
            arr[t].dlogits_bstride = g.stride(0) if g is not None else 0
            arr[t].channels = c1 - c0
        loss = torch.zeros(T, device=dev, dtype=torch.float32)
        count = torch.zeros(T + 1, device=dev, dtype=torch.int32)  # [T] = bad-label flag
        d = L.HgCeDesc(T, B, HW, ignore_index, 1.0)
        with torch.cuda.device(dev):
            L.call("hg_ce_multi", C.byref(d), arr, L.ptr(loss), L.ptr(count), C.c_void_p(count.data_ptr() + 4 * T),
                   L.stream_ptr())
        if check and int(count[T]) != 0:
            raise IndexError("cross_entropy_losses: a label is outside [0, C) and is not ignore_index")
        ctx.grads, ctx.terms = grads, [(k, c0, c1) for k, c0, c1, _ in terms]
        return loss

    @staticmethod
    def backward(ctx, gloss):
        out = [None]
        for i, g in enumerate(ctx.grads):
            if g is not None:
                for t, (k, c0, c1) in enumerate(ctx.terms):
                    if k == i:
                        g[:, c0:c1].mul_(gloss[t])
            out.append(g)
        ctx.grads = None
        return tuple(out)


def cross_entropy_losses(terms, ignore_index=-100, check_labels=False):
    """All `nn.CrossEntropyLoss` heads of a training step in one launch.  terms: list of `(logits, target)` or
    `(logits, target, (c0, c1))`; logits fp32 [B,C,64,64] (a model output), target int64 [B,64,64], (c0, c1) a channel
    slice of logits.  Returns a tensor with one mean-reduced loss per term, e.g. the objective of
    try_skeleton_and_keypoints.py:423-435:

        terms = [(r, by_keypoints, (0, 18)) for r in result] + [(r, by_skeleton, (18, 38)) for r in result]
        cross_entropy_losses(terms).sum().backward()

    Passing the slice explicitly lets every term of one output write into the same gradient tensor (no
    zero-padded slice gradients); `(result[k][:, :18], target)` on views works too.
    check_labels=True synchronises and raises IndexError on labels outside [0, C) (PyTorch's device assert)."""
    uniq, spec = [], []
    for term in terms:
        x, target = term[0], term[1]
        c0, c1 = term[2] if len(term) > 2 else (0, x.shape[1])
        for k, u in enumerate(uniq):
            if u is x:
                break
        else:
            uniq.append(x)
            k = len(uniq) - 1
        spec.append((k, int(c0), int(c1), target))
    return _CeMulti.apply((spec, int(ignore_index), bool(check_labels)), *uniq)


def _ce_call(x, tgt, nll_out=None, weight=None, norm=0.0, grad=None):
    """One hg_ce_multi term over all channels of x; returns the [1] loss tensor."""
    B, Cx, H, W = x.shape
    arr = (L.HgCeTerm * 1)()
    arr[0].logits, arr[0].target = x.data_ptr(), tgt.data_ptr()
    arr[0].dlogits = grad.data_ptr() if grad is not None else None
    arr[0].logits_bstride = x.stride(0)
    arr[0].dlogits_bstride = grad.stride(0) if grad is not None else 0
    arr[0].channels = Cx
    arr[0].norm = float(norm)
    arr[0].pixel_weight = weight.data_ptr() if weight is not None else None
    arr[0].nll_out = nll_out.data_ptr() if nll_out is not None else None
    loss = torch.zeros(1, device=x.device, dtype=torch.float32)
    count = torch.zeros(2, device=x.device, dtype=torch.int32)
    d = L.HgCeDesc(1, B, H * W, -100, 1.0)
    L.call("hg_ce_multi", C.byref(d), arr, L.ptr(loss), L.ptr(count), C.c_void_p(count.data_ptr() + 4), L.stream_ptr())
    return loss


def _check_ce_inputs(input, target, who):
    if not input.is_cuda or input.dtype != torch.float32 or input.dim() != 4:
        raise RuntimeError(f"{who}: fp32 CUDA logits [B,C,H,W] expected (there is no CPU fallback)")
    if not target.is_cuda or tuple(target.shape) != (input.shape[0],) + tuple(input.shape[2:]):
        raise RuntimeError(f"{who}: labels must be a CUDA tensor [B,H,W]")
    x = input if _plane_view(input) else input.contiguous()
    return x, target.contiguous().to(torch.int64)


class _CeWeighted(torch.autograd.Function):
    """mean-of-top-k (mask=None) or mask-weighted per-pixel cross entropy."""

    @staticmethod
    def forward(ctx, input, target, k, mask):
        x, tgt = _check_ce_inputs(input, target, "bootstrapped / masked cross entropy")
        B, Cx, H, W = x.shape
        grad = torch.empty(x.shape, device=x.device, dtype=torch.float32) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(x.device):
            if mask is None:
                nll = torch.empty(B, H * W, device=x.device, dtype=torch.float32)
                _ce_call(x, tgt, nll_out=nll, norm=1.0)
                sel = torch.empty_like(nll)
                L.call("hg_topk_mask", L.ptr(nll), B, H * W, int(k), L.ptr(sel), None, L.stream_ptr())
                loss = _ce_call(x, tgt, weight=sel, norm=float(B * k), grad=grad)
            else:
                w = mask.contiguous().to(torch.float32)
                loss = _ce_call(x, tgt, weight=w, norm=float(B * H * W), grad=grad)
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, gloss):
        g, ctx.grad = ctx.grad, None
        return (g * gloss if g is not None else None), None, None, None


def bootstrapped_cross_entropy(input, target, fraction):
    """Costomer_CrossEntropyLoss.forward of train.py:350-362: per-pixel NLL of log_softmax(input, dim=1), the
    k = int(H * W * max(fraction, 0.1)) largest of every image, their mean."""
    if fraction < 0.1:
        fraction = 0.1
    k = int(input.shape[2] * input.shape[3] * fraction)
    return _CeWeighted.apply(input, target, k, None)


def masked_cross_entropy(input, target, mask):
    """Costomer_CrossEntropyLoss_with_mask.forward of train.py:372-376: mean over ALL pixels of nll * mask."""
    return _CeWeighted.apply(input, target, 0, mask)


class _MseWeighted(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, target, k, mask):
        if not input.is_cuda or input.dtype != torch.float32 or input.dim() != 4 or input.shape != target.shape:
            raise RuntimeError("bootstrapped / masked MSE: fp32 CUDA tensors [B,C,H,W] of one shape expected "
                               "(there is no CPU fallback)")
        x, t = input.contiguous(), target.contiguous().to(torch.float32)
        B, Cx, H, W = x.shape
        HW = H * W
        grad = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        loss = torch.zeros(1, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            if mask is None:   # mean of the k largest squared errors of every image (over C*H*W elements)
                sq = torch.empty(B, Cx * HW, device=x.device, dtype=torch.float32)
                L.call("hg_mse_weighted", L.ptr(x), L.ptr(t), None, 0, B, Cx, HW, C.c_float(1.0), C.c_float(1.0),
                       L.ptr(sq), None, None, L.stream_ptr())
                sel = torch.empty_like(sq)
                L.call("hg_topk_mask", L.ptr(sq), B, Cx * HW, int(k), L.ptr(sel), None, L.stream_ptr())
                L.call("hg_mse_weighted", L.ptr(x), L.ptr(t), L.ptr(sel), 0, B, Cx, HW, C.c_float(float(B * k)),
                       C.c_float(1.0), None, L.ptr(grad), L.ptr(loss), L.stream_ptr())
            else:
                w = mask.contiguous().to(torch.float32)
                L.call("hg_mse_weighted", L.ptr(x), L.ptr(t), L.ptr(w), 1, B, Cx, HW, C.c_float(float(x.numel())),
                       C.c_float(1.0), None, L.ptr(grad), L.ptr(loss), L.stream_ptr())
        ctx.grad = grad
        return loss[0]

    @staticmethod
    def backward(ctx, gloss):
        g, ctx.grad = ctx.grad, None
        return (g * gloss if g is not None else None), None, None, None


def bootstrapped_mse(input, target, fraction):
    """Costomer_MSELoss.forward of train.py:401-408: squared errors, the k = int(H * W * max(fraction, 0.25)) largest
    of every image (over all C*H*W elements), their mean."""
    if fraction < 0.25:
        fraction = 0.25
    k = int(input.shape[2] * input.shape[3] * fraction)
    return _MseWeighted.apply(input, target, k, None)


def masked_mse(input, target, mask):
    """Costomer_MSELoss_with_mask.forward of train.py:386-391: mean over all elements of (input - target)^2 * mask,
    mask [B,H,W] broadcast over the channels."""
    return _MseWeighted.apply(input, target, 0, mask)
