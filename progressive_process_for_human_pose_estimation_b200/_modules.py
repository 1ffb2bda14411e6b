"""Host-side mirror of the reference's model-construction API.

The reference is a set of flat scripts that each re-declare `ResidualBlock`, `hourglass`, `lin`, `creatModel`
and read their configuration (`nStack`, `nModules`, `nFeats`, `nOutChannels`) from *module globals at call
time* (try_with_torch.py:23-33,224,285).  The mirror modules of this package (try_with_torch.py,
only_one_hourgless.py, ...) keep exactly that surface: same class names, constructor signatures, sub-module
names (hence `state_dict` keys, Appendix A of SURVEY.md) and list-of-heatmaps return value.  The classes are
built by the factories below with a reference to the mirror module's `globals()` so that a caller can still do
`mod.nStack = 8` before constructing / calling a model.

Sub-modules are stock `nn.Conv2d` / `nn.BatchNorm2d` objects created in the reference's order (seeded
initialisation and checkpoints are therefore interchangeable with the reference), but they are only parameter
containers: `forward` never calls them.  Instead every class has `_emit(builder, x)` which replays the
reference's forward as graph ops; `forward(x)` looks up / builds the execution plan for (shape, mode, config)
and runs it through the C-ABI library (plan.py).  There is no PyTorch fallback.
"""
import torch
import torch.nn as nn

from . import _lib as L
from .plan import Builder, Plan, run_plan

_compute_dtype = torch.bfloat16


def set_compute_dtype(dtype):
    """torch.bfloat16 (tensor-core path, default) or torch.float32 (CUDA-core path, rtol 1e-5)."""
    global _compute_dtype
    if dtype not in (torch.bfloat16, torch.float32):
        raise ValueError("compute dtype must be torch.bfloat16 or torch.float32")
    _compute_dtype = dtype


def get_compute_dtype():
    return _compute_dtype


class HGModule(nn.Module):
    """Base class: plan cache + dispatch.  Subclasses implement `_emit(b, x) -> Val` (or a list for models)."""

    _is_model = False  # models take the fp32 NCHW image batch and return a list of heatmaps

    def _config_key(self):
        return ()

    def _plan_for(self, x):
        if not x.is_cuda:
            raise RuntimeError(
                f"{type(self).__name__}: the hourglass hot path runs only on a CUDA (sm_100a) device; there is no "
                "CPU fallback. Move the module and its input with .cuda().")
        if x.dim() != 4:
            raise RuntimeError(f"{type(self).__name__}: expected a 4-D NCHW input, got shape {tuple(x.shape)}")
        params = list(self.named_parameters())
        train_params = torch.is_grad_enabled() and any(p.requires_grad for _, p in params)
        x_rg = torch.is_grad_enabled() and x.requires_grad
        key = (tuple(x.shape), self.training, train_params, x_rg, _compute_dtype, self._config_key(),
               tuple(p.data_ptr() for _, p in params))
        cache = self.__dict__.setdefault("_plans", {})
        plan = cache.get(key)
        if plan is None:
            L.load()
            for _, p in params:
                if p.device != x.device:
                    raise RuntimeError(f"{type(self).__name__}: parameters and input live on different devices")
            if self.training:
                for name, buf in self.named_buffers():
                    if buf.is_floating_point() and buf.dtype != torch.float32:
                        raise RuntimeError("training-mode BatchNorm needs fp32 running statistics "
                                           f"(buffer {name} is {buf.dtype})")
            b = Builder(self.training, train_params)
            if self._is_model:
                if x_rg:
                    raise RuntimeError("gradients with respect to the input image are not part of the hot path")
                xin = b.input_image(x.shape[0], x.shape[2], x.shape[3])
                if x.shape[1] != 3:
                    raise RuntimeError(f"{type(self).__name__}: expected a 3-channel image batch")
                self._emit(b, xin)
            else:
                xin = b.input_nchw(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x_rg)
                out = self._emit(b, xin)
                b.output(out)
            with torch.cuda.device(x.device):
                plan = Plan(b, params, x.device, _compute_dtype)
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[key] = plan
        return plan

    def forward(self, x):
        plan = self._plan_for(x)
        with torch.cuda.device(x.device):
            outs = run_plan(plan, x)
        if x.dtype != torch.float32:
            outs = [o.to(x.dtype) for o in outs]
        return outs if self._is_model else outs[0]

    def __getstate__(self):
        st = super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__.copy()
        st = dict(st)
        st.pop("_plans", None)
        return st

    def launches_per_step(self):
        """(forward, backward) C-ABI kernel launches of the most recently built plan."""
        plans = self.__dict__.get("_plans", {})
        if not plans:
            return (0, 0)
        p = list(plans.values())[-1]
        return (p.launches_fwd, p.launches_bwd)


def make_s_family(g):
    """Classes of the weight-shared recursive family (try_with_torch.py:179-298, only_one_hourgless.py:135-254,
    try_with_torch_100.py:117-252).  `g` is the mirror module's globals()."""

    class ResidualBlock(HGModule):
        """Pre-activation bottleneck, biased convs, 1x1 projection only when numIn != numOut
        (try_with_torch.py:179-209).  `conv4` exists even when unused, as in the reference."""

        def __init__(self, numIn, numOut):
            super(ResidualBlock, self).__init__()
            self.numIn = numIn
            self.numOut = numOut
            self.bn1 = nn.BatchNorm2d(numIn)
            self.relu = nn.ReLU(True)
            self.conv1 = nn.Conv2d(numIn, int(numOut / 2), 1, 1)
            self.bn2 = nn.BatchNorm2d(int(numOut / 2))
            self.relu = nn.ReLU(True)
            self.conv2 = nn.Conv2d(int(numOut / 2), int(numOut / 2), 3, 1, 1)
            self.bn3 = nn.BatchNorm2d(int(numOut / 2))
            self.relu = nn.ReLU(True)
            self.conv3 = nn.Conv2d(int(numOut / 2), numOut, 1, 1)
            self.conv4 = nn.Conv2d(numIn, numOut, 1, 1)

        def _emit(self, b, x):
            a1 = b.bn_relu(self.bn1, x)
            y1 = b.conv(self.conv1, a1)
            a2 = b.bn_relu(self.bn2, y1)
            y2 = b.conv(self.conv2, a2)
            a3 = b.bn_relu(self.bn3, y2)
            residual = x if self.numIn == self.numOut else b.conv(self.conv4, x)
            return b.conv(self.conv3, a3, residual=residual)  # `out += residual` fused into the epilogue

    class hourglass(HGModule):
        """Recursive hourglass whose levels each own ONE residual block applied 6-8 times (quirk Q1),
        bilinear align_corners=True up-sampling fused with the skip add (try_with_torch.py:212-240)."""

        def __init__(self, n, f):
            super(hourglass, self).__init__()
            self.n = n
            self.f = f
            self.residual_block = ResidualBlock(f, f)
            if n > 1:
                self.hourglass1 = hourglass(n - 1, f)
            self.maxpool = nn.MaxPool2d(2)

        def _config_key(self):
            return (g["nModules"],)

        def _emit(self, b, x):
            nModules = g["nModules"]
            up1 = x
            # the skip branch is independent of the whole low-resolution path: it gets its own stream lane so the
            # launch-latency-bound 4x4 .. 16x16 chain overlaps the large high-resolution kernels
            with b.on_lane(self.n):
                for _ in range(nModules):
                    up1 = self.residual_block._emit(b, up1)
            low1 = b.maxpool2(x)
            for _ in range(nModules):
                low1 = self.residual_block._emit(b, low1)
            if self.n > 1:
                low2 = self.hourglass1._emit(b, low1)
            else:
                low2 = low1
                for _ in range(nModules):
                    low2 = self.residual_block._emit(b, low2)
            low3 = low2
            for _ in range(nModules):
                low3 = self.residual_block._emit(b, low3)
            return b.upsample2x_add(low3, up1, mode="bilinear")

    class lin(HGModule):
        """1x1 conv + BN + ReLU (try_with_torch.py:243-256)."""

        def __init__(self, numIn, numOut):
            super(lin, self).__init__()
            self.numIn = numIn
            self.numOut = numOut
            self.conv = nn.Conv2d(numIn, numOut, 1, 1, 0)
            self.bn = nn.BatchNorm2d(numOut)
            self.relu = nn.ReLU()

        def _emit(self, b, x):
            return b.bn_relu(self.bn, b.conv(self.conv, x))

    class creatModel(HGModule):
        """N-stack weight-shared hourglass network with one heatmap head reused by every stack
        (try_with_torch.py:259-298).  forward(x[B,3,256,256]) -> list of nStack tensors [B,nOutChannels,64,64]."""

        _is_model = True

        def __init__(self):
            super(creatModel, self).__init__()
            nFeats, nOutChannels = g["nFeats"], g["nOutChannels"]
            self.conv1 = nn.Conv2d(3, 64, 7, 2, 3)
            self.relu = nn.ReLU()
            self.residual1 = ResidualBlock(64, 128)
            self.max_pool1 = nn.MaxPool2d(2)
            self.residual2 = ResidualBlock(128, 128)
            self.residual3 = ResidualBlock(128, nFeats)
            self.hourglass1 = hourglass(4, nFeats)
            self.residual4 = ResidualBlock(nFeats, nFeats)
            self.lin = lin(nFeats, nFeats)
            self.conv2 = nn.Conv2d(nFeats, nOutChannels, 1, 1, 0)
            self.conv3 = nn.Conv2d(nFeats, nFeats, 1, 1, 0)
            self.conv4 = nn.Conv2d(nOutChannels, nFeats, 1, 1, 0)

        def _config_key(self):
            return (g["nStack"], g["nModules"])

        def _emit(self, b, x):
            nStack, nModules = g["nStack"], g["nModules"]
            x = b.stem(self.conv1, x)
            x = self.residual1._emit(b, x)
            x = b.maxpool2(x)
            x = self.residual2._emit(b, x)
            x = self.residual3._emit(b, x)
            out = []
            inter = x
            for i in range(nStack):
                hg = self.hourglass1._emit(b, inter)
                ll = hg
                for _ in range(nModules):
                    ll = self.residual4._emit(b, ll)
                ll = self.lin._emit(b, ll)
                out_keypoints = b.conv(self.conv2, ll, head=True)
                out.insert(i, out_keypoints)
                if i < nStack:  # always true, as in the reference (quirk Q5)
                    ll_ = b.conv(self.conv3, ll)
                    inter = b.conv(self.conv4, out_keypoints, residual=ll_)
            return out

    for cls in (ResidualBlock, hourglass, lin, creatModel):
        cls.__module__ = g.get("__name__", cls.__module__)
        cls.__qualname__ = cls.__name__
    return ResidualBlock, hourglass, lin, creatModel
