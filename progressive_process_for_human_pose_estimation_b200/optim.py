"""Adam in one kernel launch (csrc/optim.cu), drop-in for `torch.optim.Adam(model.parameters(), lr=1e-4)` of the
reference training loops (try_with_torch.py:317,342-344; eps=1e-4 in hourglass_compare.py:885).

The optimizer state keeps torch.optim.Adam's layout ('step', 'exp_avg', 'exp_avg_sq' per parameter, same param_groups
keys), so `opt.load_state_dict(state['optimizer'])` of a reference checkpoint (try_with_torch.py:324-328) and
`torch.save({'optimizer': opt.state_dict()})` (:361-367) interoperate with the stock class.  fp32 CUDA parameters
only; there is no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib as L

_CHUNK = 16384  # elements per thread block


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise RuntimeError("hg.Adam: amsgrad is not implemented (the reference never enables it)")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("hg.Adam: invalid hyper-parameters")
        # the extra keys are torch.optim.Adam's own defaults: param_groups stay loadable by the stock optimizer
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        super().__init__(params, defaults)
        self._tables = {}   # group index -> (key, device chunk table, number of chunks)

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in plist)
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        chunks = []
        for p in plist:
            st, n = self.state[p], p.numel()
            for o in range(0, n, _CHUNK):
                chunks.append(L.HgAdamChunk(p.data_ptr() + 4 * o, p.grad.data_ptr() + 4 * o,
                                            st["exp_avg"].data_ptr() + 4 * o, st["exp_avg_sq"].data_ptr() + 4 * o,
                                            min(_CHUNK, n - o)))
        arr = (L.HgAdamChunk * len(chunks))(*chunks)
        dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(plist[0].device)
        self._tables[gi] = (key, dev, len(chunks))
        return dev, len(chunks)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            steps = set()
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("hg.Adam: fp32 CUDA parameters and gradients only (there is no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("hg.Adam: parameters and gradients must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                steps.add(float(st["step"]))
            if len(steps) != 1:   # parameters that joined later: one launch per distinct step count
                groups = {}
                for p in plist:
                    groups.setdefault(float(self.state[p]["step"]), []).append(p)
            else:
                groups = {steps.pop(): plist}
            for si, (step0, ps) in enumerate(sorted(groups.items())):
                table, n = self._table((gi, si, len(groups)), ps)
                b1, b2 = group["betas"]
                d = L.HgAdamDesc(float(group["lr"]), float(b1), float(b2), float(group["lr"]), float(b1), float(b2),
                                 float(group["eps"]), float(group["weight_decay"]), int(step0) + 1, n, 0)
                with torch.cuda.device(ps[0].device):
                    L.call("hg_adam_multi", C.byref(d), L.ptr(table), L.stream_ptr())
                for p in ps:
                    self.state[p]["step"] += 1
        return loss
