"""Fused intermediate-supervision MSE loss (csrc/loss.cu).

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
        loss = torch.zeros(S, device=target.device, dtype=torch.float32)
        d = L.HgMseDesc(target.numel(), S, 1.0)
        parr = (C.c_void_p * S)(*[p.data_ptr() for p in preds])
        garr = (C.c_void_p * S)(*[g.data_ptr() if g is not None else None for g in grads])
        with torch.cuda.device(target.device):
            L.call("hg_mse_multi", C.byref(d), parr, L.ptr(target), garr, L.ptr(loss), L.stream_ptr())
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, gloss):
        out = [None]
        for s, g in enumerate(ctx.grads):
            out.append(None if g is None else g * gloss[s])
        ctx.grads = None
        return tuple(out)


def mse_losses(result, target):
    """Per-stack MSE losses as one tensor [nStack]; `mse_losses(result, y).sum().backward()` is the training
    objective of try_with_torch.py:333-342."""
    return _MseMulti.apply(target, *result)
