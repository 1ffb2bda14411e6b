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
    _single_output = False  # a model whose reference forward returns the head's tensor itself (train.generateMask)

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
        # lowering bakes in every parameter's requires_grad (gradient slots, frozen-parameter pruning) and every
        # BatchNorm's own train / eval flag, so both are part of the key: unfreezing a parameter or calling .eval() on a
        # sub-module after the first forward builds a new plan, as a stock nn.Module would simply honour the change
        bn_flags = tuple(mod.training for mod in self.modules() if isinstance(mod, nn.BatchNorm2d))
        key = (tuple(x.shape), self.training, train_params, x_rg, _compute_dtype, self._config_key(),
               tuple(p.data_ptr() for _, p in params),
               tuple(p.requires_grad for _, p in params) if torch.is_grad_enabled() else (), bn_flags)
        cache = self.__dict__.setdefault("_plans", {})
        plan = cache.get(key)
        if plan is None:
            L.load()
            for _, p in params:
                if p.device != x.device:
                    raise RuntimeError(f"{type(self).__name__}: parameters and input live on different devices")
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
        return outs if (self._is_model and not self._single_output) else outs[0]

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


def make_aspp_block(g=None):
    """`_ASPPModule` (try_with_aspp.py:193-205 = train.py:449-461: dilated conv, no bias -> BN -> ReLU) and `ASPP_Block`
    (train.py:465-495: 1x1 + three dilated 3x3 branches (6 / 12 / 18) + image-level branch (global average pool ->
    1x1 -> BN -> ReLU -> broadcast), concatenated (1280 ch) -> 1x1 -> BN -> ReLU) as executable drop-in modules."""

    class _ASPPModule(HGModule):
        def __init__(self, inplanes, planes, kernel_size, padding, dilation):
            super(_ASPPModule, self).__init__()
            self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size=kernel_size, stride=1, padding=padding,
                                         dilation=dilation, bias=False)
            self.bn = nn.BatchNorm2d(planes)
            self.relu = nn.ReLU()

        def _emit(self, b, x):
            return b.bn_relu(self.bn, b.conv(self.atrous_conv, x))

    class ASPP_Block(HGModule):
        def __init__(self):
            super(ASPP_Block, self).__init__()
            inplanes = 256
            dilations = [1, 6, 12, 18]
            self.aspp1 = _ASPPModule(inplanes, 256, 1, padding=0, dilation=dilations[0])
            self.aspp2 = _ASPPModule(inplanes, 256, 3, padding=dilations[1], dilation=dilations[1])
            self.aspp3 = _ASPPModule(inplanes, 256, 3, padding=dilations[2], dilation=dilations[2])
            self.aspp4 = _ASPPModule(inplanes, 256, 3, padding=dilations[3], dilation=dilations[3])
            self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)),
                                                 nn.Conv2d(inplanes, 256, 1, stride=1, bias=False),
                                                 nn.BatchNorm2d(256),
                                                 nn.ReLU())
            self.conv1 = nn.Sequential(nn.Conv2d(1280, 256, 1, bias=False), nn.BatchNorm2d(256), nn.ReLU())

        def _emit(self, b, x):
            # the five branches only share their input: one stream lane each
            xs = []
            for i, m in enumerate((self.aspp1, self.aspp2, self.aspp3, self.aspp4)):
                with b.on_lane(i):
                    xs.append(m._emit(b, x))
            with b.on_lane(4):
                x5 = b.global_avg_pool(x)
                x5 = b.bn_relu(self.global_avg_pool[2], b.conv(self.global_avg_pool[1], x5))
                xs.append(b.broadcast_to(x5, x.H, x.W))
            # torch.cat + 1x1 conv as five chained convolutions over slices of the 1280-channel weight
            return b.bn_relu(self.conv1[1], b.conv_cat(self.conv1[0], xs))

    for cls in (_ASPPModule, ASPP_Block):
        if g is not None:
            cls.__module__ = g.get("__name__", cls.__module__)
        cls.__qualname__ = cls.__name__
    return _ASPPModule, ASPP_Block


def make_train_family(g):
    """train.py:411-601, the progressive multi-branch model: Q4 residual blocks, an un-shared hourglass
    whose four down-sampling steps are stride-2 blocks, whose bottom is the executed ASPP block and whose up path is
    nearest x2 -> block(f -> f/2) -> cat with the skip branch; three stages with bias-free heads (2 / 16 / 17 channels
    by default) re-injected as cat[conv(head), conv(ll), conv(inter)]."""
    ResidualBlock = make_q4_block(g)
    _ASPPModule, ASPP_Block = make_aspp_block(g)

    class hourglass(HGModule):
        def __init__(self, f):
            super(hourglass, self).__init__()
            self.f = f
            self.downsample1 = ResidualBlock(f, f, stride=2)
            self.downsample2 = ResidualBlock(f, f, stride=2)
            self.downsample3 = ResidualBlock(f, f, stride=2)
            self.downsample4 = ResidualBlock(f, f, stride=2)
            self.residual1 = ResidualBlock(f, int(f / 2))
            self.residual2 = ResidualBlock(f, int(f / 2))
            self.residual3 = ResidualBlock(f, int(f / 2))
            self.residual4 = ResidualBlock(f, int(f / 2))
            self.upsample1 = ResidualBlock(f, int(f / 2))
            self.upsample2 = ResidualBlock(f, int(f / 2))
            self.upsample3 = ResidualBlock(f, int(f / 2))
            self.upsample4 = ResidualBlock(f, int(f / 2))
            self.aspp = ASPP_Block()

        def _emit(self, b, x):
            ups, down = [], x
            for lvl, (res, ds) in enumerate(((self.residual1, self.downsample1), (self.residual2, self.downsample2),
                                             (self.residual3, self.downsample3), (self.residual4, self.downsample4))):
                with b.on_lane(lvl + 1):           # skip branches are independent of the deeper path
                    ups.append(res._emit(b, down))
                down = ds._emit(b, down)
            out = self.aspp._emit(b, down)
            for up_block, skip in ((self.upsample4, ups[3]), (self.upsample3, ups[2]), (self.upsample2, ups[1]),
                                   (self.upsample1, ups[0])):
                out = b.upsample2x_add(out, None, mode="nearest")   # F.interpolate(out, scale_factor=2)
                out = up_block._emit(b, out)
                out = b.cat([out, skip])
            return out

    class creatModel(HGModule):
        _is_model = True

        def __init__(self):
            super(creatModel, self).__init__()
            nFeats = g["nFeats"]
            self.preprocess1 = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3), nn.ReLU(), ResidualBlock(64, 128, stride=2),
                                             ResidualBlock(128, 128), ResidualBlock(128, nFeats))
            self.stage1 = hourglass(nFeats)
            self.stage1_out = nn.Conv2d(nFeats, g["nOutChannels_0"], 1, 1, 0, bias=False)
            self.stage1_return = nn.Conv2d(g["nOutChannels_0"], int(nFeats / 2), 1, 1, 0, bias=False)
            self.stage1_retuen_2 = nn.Conv2d(nFeats, int(nFeats / 4), 1, 1, 0, bias=False)
            self.stage1_down_feature = nn.Conv2d(nFeats, int(nFeats / 4), 1, 1, 0, bias=False)
            self.stage2 = hourglass(nFeats)
            self.stage2_out = nn.Conv2d(nFeats, g["nOutChannels_1"], 1, 1, 0, bias=False)
            self.stage2_return = nn.Conv2d(g["nOutChannels_1"], int(nFeats / 2), 1, 1, 0, bias=False)
            self.stage2_retuen_2 = nn.Conv2d(nFeats, int(nFeats / 4), 1, 1, 0, bias=False)
            self.stage2_down_feature = nn.Conv2d(nFeats, int(nFeats / 4), 1, 1, 0, bias=False)
            self.stage3 = hourglass(nFeats)
            self.stage3_out = nn.Conv2d(nFeats, g["nOutChannels_2"], 1, 1, 0, bias=False)

        def _emit(self, b, x):
            pre = self.preprocess1
            inter = b.stem(pre[0], x)
            for blk in (pre[2], pre[3], pre[4]):
                inter = blk._emit(b, inter)
            out = []
            for i, (stage, head, ret, ret2, down) in enumerate((
                    (self.stage1, self.stage1_out, self.stage1_return, self.stage1_retuen_2, self.stage1_down_feature),
                    (self.stage2, self.stage2_out, self.stage2_return, self.stage2_retuen_2, self.stage2_down_feature),
                    (self.stage3, self.stage3_out, None, None, None))):
                ll = stage._emit(b, inter)
                tmpOut = b.conv(head, ll, head=True)
                out.insert(i, tmpOut)
                if ret is not None:
                    inter = b.cat([b.conv(ret, tmpOut), b.conv(ret2, ll), b.conv(down, inter)])
            return out

    class generateMask(HGModule):
        """train.py:604-622: the first stage alone (stem, three blocks, one hourglass, the 2-channel head); returns the
        head's tensor itself, not a list."""

        _is_model = True
        _single_output = True

        def __init__(self):
            super(generateMask, self).__init__()
            nFeats = g["nFeats"]
            self.preprocess1 = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3), nn.ReLU(), ResidualBlock(64, 128, stride=2),
                                             ResidualBlock(128, 128), ResidualBlock(128, nFeats))
            self.stage1 = hourglass(nFeats)
            self.stage1_out = nn.Conv2d(nFeats, g["nOutChannels_0"], 1, 1, 0, bias=False)

        def _emit(self, b, x):
            pre = self.preprocess1
            inter = b.stem(pre[0], x)
            for blk in (pre[2], pre[3], pre[4]):
                inter = blk._emit(b, inter)
            return [b.conv(self.stage1_out, self.stage1._emit(b, inter), head=True)]

    for cls in (hourglass, creatModel, generateMask):
        cls.__module__ = g.get("__name__", cls.__module__)
        cls.__qualname__ = cls.__name__
    return ResidualBlock, _ASPPModule, ASPP_Block, hourglass, creatModel, generateMask


def _multihead_forward(self, b, x, g, cat_inter, with_pool, n_res4, repeat_last_head=False):
    """Shared body of the multi-head creatModel variants: per-stack heads conv2_k and re-injection conv4_k over a
    concatenation (try_different_stack.py:300-329; try_with_aspp_remove_max_pool.py:277-304)."""
    nStack = g["nStack"]
    x = b.stem(self.conv1, x)
    x = self.residual1._emit(b, x)
    if with_pool:
        x = b.maxpool2(x)
    x = self.residual2._emit(b, x)
    x = self.residual3._emit(b, x)
    out = []
    inter = x
    heads = [getattr(self, f"conv2_{k}") for k in range(3) if hasattr(self, f"conv2_{k}")]
    reinj = [getattr(self, f"conv4_{k}") for k in range(2) if hasattr(self, f"conv4_{k}")]
    for i in range(nStack):
        hg = self.hourglass1._emit(b, inter)
        ll = hg
        for _ in range(n_res4):
            ll = self.residual4._emit(b, ll)
        ll = self.lin._emit(b, ll)
        if i >= len(heads):
            if not repeat_last_head:
                continue  # the reference has no branch for further stacks: nothing is appended
            # try_more_layer.py:339-341 (`elif i >= 2`): every further stack goes through the last head, no re-injection
            out.insert(i, b.conv(heads[-1], ll, head=True))
            continue
        tmpOut = b.conv(heads[i], ll, head=True)
        out.insert(i, tmpOut)
        if i < len(reinj):
            parts = [inter, ll, tmpOut] if cat_inter else [ll, tmpOut]
            inter = b.conv_cat(reinj[i], parts)
    return out


def make_multihead_family(g, aspp_members=False, num_heads=3, aspp_executed=False):
    """try_different_stack.py / try_different_stack_without_skeleton.py (aspp_members=False) and try_with_aspp.py
    (aspp_members=True: ASPP modules are constructed -- they are in the state_dict -- but never executed, and the
    bottom level has no extra residual blocks, try_with_aspp.py:250-279).  aspp_executed=True is try_more_layer.py
    (:249-296,339-341): the bottom level RUNS the inline ASPP (four dilated branches + image-level branch -> cat ->
    1x1 conv without BN), and every stack beyond the third reuses the last head."""
    ResidualBlock, hourglass_s, lin, _ = make_s_family(g)

    # never called by this family's forward (quirk Q6) but executable on its own, like the reference class
    _ASPPModule, _ = make_aspp_block(g)

    if aspp_members:
        class hourglass(HGModule):
            def __init__(self, n, f):
                super(hourglass, self).__init__()
                self.n = n
                self.f = f
                self.residual_block = ResidualBlock(f, f)
                if n > 1:
                    self.hourglass1 = hourglass(n - 1, f)
                self.maxpool = nn.MaxPool2d(2)
                inplanes = 256
                dilations = [1, 6, 12, 18]
                self.aspp1 = _ASPPModule(inplanes, 256, 1, padding=0, dilation=dilations[0])
                self.aspp2 = _ASPPModule(inplanes, 256, 3, padding=dilations[1], dilation=dilations[1])
                self.aspp3 = _ASPPModule(inplanes, 256, 3, padding=dilations[2], dilation=dilations[2])
                self.aspp4 = _ASPPModule(inplanes, 256, 3, padding=dilations[3], dilation=dilations[3])
                self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)),
                                                     nn.Conv2d(inplanes, 256, 1, stride=1, bias=False),
                                                     nn.BatchNorm2d(256), nn.ReLU())
                self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)

            def _config_key(self):
                return (g["nModules"],)

            def _emit(self, b, x):
                nModules = g["nModules"]
                up1 = x
                with b.on_lane(self.n):
                    for _ in range(nModules):
                        up1 = self.residual_block._emit(b, up1)
                low1 = b.maxpool2(x)
                for _ in range(nModules):
                    low1 = self.residual_block._emit(b, low1)
                if self.n > 1:
                    low2 = self.hourglass1._emit(b, low1)
                elif aspp_executed:   # try_more_layer.py:281-290
                    xs = [m._emit(b, low1) for m in (self.aspp1, self.aspp2, self.aspp3, self.aspp4)]
                    x5 = b.global_avg_pool(low1)
                    x5 = b.bn_relu(self.global_avg_pool[2], b.conv(self.global_avg_pool[1], x5))
                    xs.append(b.broadcast_to(x5, low1.H, low1.W))
                    low2 = b.conv_cat(self.conv1, xs)
                else:
                    low2 = low1
                low3 = low2
                for _ in range(nModules):
                    low3 = self.residual_block._emit(b, low3)
                return b.upsample2x_add(low3, up1, mode="bilinear")
    else:
        hourglass = hourglass_s

    class creatModel(HGModule):
        """3-stack network with a different head per stack: 2-ch background, 20-ch limb, 17-ch keypoint maps
        (try_different_stack.py:282-329)."""

        _is_model = True

        def __init__(self):
            super(creatModel, self).__init__()
            nFeats = g["nFeats"]
            self.conv1 = nn.Conv2d(3, 64, 7, 2, 3)
            self.relu = nn.ReLU()
            self.residual1 = ResidualBlock(64, 128)
            self.max_pool1 = nn.MaxPool2d(2)
            self.residual2 = ResidualBlock(128, 128)
            self.residual3 = ResidualBlock(128, nFeats)
            self.hourglass1 = hourglass(4, nFeats)
            self.residual4 = ResidualBlock(nFeats, nFeats)
            self.lin = lin(nFeats, nFeats)
            self.conv2_0 = nn.Conv2d(nFeats, g["nOutChannels_0"], 1, 1, 0, bias=False)
            self.conv4_0 = nn.Conv2d(nFeats + g["nOutChannels_0"], nFeats, 1, 1, 0)
            self.conv2_1 = nn.Conv2d(nFeats, g["nOutChannels_1"], 1, 1, 0, bias=False)
            if num_heads >= 3:  # try_different_stack_without_skeleton.py:294-297 stops after conv2_1
                self.conv4_1 = nn.Conv2d(nFeats + g["nOutChannels_1"], nFeats, 1, 1, 0, bias=False)
                self.conv2_2 = nn.Conv2d(nFeats, g["nOutChannels_2"], 1, 1, 0, bias=False)

        def _config_key(self):
            return (g["nStack"], g["nModules"])

        def _emit(self, b, x):
            return _multihead_forward(self, b, x, g, cat_inter=False, with_pool=True, n_res4=g["nModules"],
                                      repeat_last_head=aspp_executed)

    for cls in (hourglass, creatModel, _ASPPModule):
        cls.__module__ = g.get("__name__", cls.__module__)
        cls.__qualname__ = cls.__name__
    return ResidualBlock, hourglass, lin, creatModel, _ASPPModule


def make_skeleton_family(g):
    """try_skeleton_and_keypoints.py: S-family network whose 38-channel head (18 keypoint classes + 20 limb classes)
    is mixed in place after being appended to the output list (quirk Q11, :274-300):
        t[:, 19+l] = t[:, 19+l] - t[:, 0] + t[:, sks[l][0]] + t[:, sks[l][1]],  l = 0..18.
    The mix is linear in the head output, so it is folded into the head's weights (plan.py `mix`)."""
    ResidualBlock, hourglass, lin, _ = make_s_family(g)

    def mix_matrix():
        C = g["nOutChannels"]
        sks = g["sks"]
        T = torch.eye(C)
        for l, (a, c) in enumerate(sks):
            r = 19 + l
            T[r, 0] -= 1.0
            T[r, a] += 1.0
            T[r, c] += 1.0
        return T

    class creatModel(HGModule):
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
            return (g["nStack"], g["nModules"], tuple(map(tuple, g["sks"])))

        def _emit(self, b, x):
            nStack, nModules = g["nStack"], g["nModules"]
            T = mix_matrix()
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
                tmpOut = b.conv(self.conv2, ll, head=True, mix=T)  # the list holds the MIXED tensor (in-place op)
                out.insert(i, tmpOut)
                if i < nStack:
                    ll_ = b.conv(self.conv3, ll)
                    inter = b.conv(self.conv4, tmpOut, residual=ll_)
            return out

    creatModel.__module__ = g.get("__name__", creatModel.__module__)
    creatModel.__qualname__ = "creatModel"
    return ResidualBlock, hourglass, lin, creatModel


def make_merge_family(g):
    """try_skeleton_from_keypoints_merge.py:264-305: S-family network whose 17-channel keypoint head is
    extended with 19 limb maps GATHERED from it, out_skeleton[l] = k[sks[l][0]] + k[sks[l][1]], tmpOut = cat(k, skeleton)
    (36 channels, returned to the loss and re-injected through conv4: 36 -> nFeats).  The gather-add and the cat are
    one linear map T = [I; G] of the head's output, so the head runs as ONE convolution with the folded weights
    T W / T b (plan.py `mix`, rectangular), and dW = T^T dW_eff: no gather kernel, no 36-channel copy."""
    ResidualBlock, hourglass, lin, _ = make_s_family(g)

    def gather_matrix():
        C = g["nOutChannels"]
        sks = g["sks"]
        T = torch.zeros(C + len(sks), C)
        T[:C] = torch.eye(C)
        for l, (a, c) in enumerate(sks):
            T[C + l, a] += 1.0
            T[C + l, c] += 1.0
        return T

    class creatModel(HGModule):
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
            self.conv4 = nn.Conv2d(nOutChannels + 19, nFeats, 1, 1, 0)

        def _config_key(self):
            return (g["nStack"], g["nModules"], tuple(map(tuple, g["sks"])))

        def _emit(self, b, x):
            nStack, nModules = g["nStack"], g["nModules"]
            T = gather_matrix()
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
                tmpOut = b.conv(self.conv2, ll, head=True, mix=T)   # cat(out_keypoints, out_skeleton)
                out.insert(i, tmpOut)
                if i < nStack:
                    ll_ = b.conv(self.conv3, ll)
                    inter = b.conv(self.conv4, tmpOut, residual=ll_)
            return out

    creatModel.__module__ = g.get("__name__", creatModel.__module__)
    creatModel.__qualname__ = "creatModel"
    return ResidualBlock, hourglass, lin, creatModel


def make_q4_block(g):
    """ResidualBlock of the later scripts (try_with_aspp_remove_max_pool.py:165-201 = hourglass_compare.py:405-441 =
    train.py:411-447): stride on the 3x3, bn4 after conv3, and -- because `self.stride != 1 | self.numIn !=
    self.numOut` parses as a chained comparison that is always true (quirk Q4) -- the skip is ALWAYS
    BN(conv1x1(x, stride)), even for 256->256 stride 1."""

    class ResidualBlock(HGModule):
        def __init__(self, numIn, numOut, stride=1):
            super(ResidualBlock, self).__init__()
            self.stride = stride
            self.numIn = numIn
            self.numOut = numOut
            self.bn1 = nn.BatchNorm2d(numIn)
            self.relu = nn.ReLU(True)
            self.conv1 = nn.Conv2d(numIn, int(numOut / 2), 1, 1)
            self.bn2 = nn.BatchNorm2d(int(numOut / 2))
            self.relu = nn.ReLU(True)
            self.conv2 = nn.Conv2d(int(numOut / 2), int(numOut / 2), 3, stride, 1)
            self.bn3 = nn.BatchNorm2d(int(numOut / 2))
            self.relu = nn.ReLU(True)
            self.conv3 = nn.Conv2d(int(numOut / 2), numOut, 1, 1)
            self.bn4 = nn.BatchNorm2d(numOut)
            self.downsaple = nn.Sequential(nn.Conv2d(numIn, numOut, 1, stride=stride, bias=False),
                                           nn.BatchNorm2d(numOut))

        def _emit(self, b, x):
            a1 = b.bn_relu(self.bn1, x)
            y1 = b.conv(self.conv1, a1)
            a2 = b.bn_relu(self.bn2, y1)
            y2 = b.conv(self.conv2, a2)
            a3 = b.bn_relu(self.bn3, y2)
            y3 = b.conv(self.conv3, a3)
            out = b.bn_relu(self.bn4, y3, relu=False)
            assert (self.stride != 1 | self.numIn != self.numOut), "quirk Q4: true for every shape the reference uses"
            r = b.conv(self.downsaple[0], x)
            r = b.bn_relu(self.downsaple[1], r, relu=False)
            return b.add(out, r)

    ResidualBlock.__module__ = g.get("__name__", ResidualBlock.__module__)
    ResidualBlock.__qualname__ = "ResidualBlock"
    return ResidualBlock


def make_nopool_family(g):
    """try_with_aspp_remove_max_pool.py (BASELINE config 4): max-pool replaced by a stride-2 block, skip merged by
    cat + 1x1 conv, unused ASPP members, multi-head creatModel with cat[inter, ll, tmpOut] re-injection."""
    ResidualBlock = make_q4_block(g)
    _, _, lin, _ = make_s_family(g)

    _ASPPModule, _ = make_aspp_block(g)

    class hourglass(HGModule):
        def __init__(self, n, f):
            super(hourglass, self).__init__()
            self.n = n
            self.f = f
            self.residual_block = ResidualBlock(f, f)
            self.residual_block_stride = ResidualBlock(f, f, stride=2)
            if n > 1:
                self.hourglass1 = hourglass(n - 1, f)
            self.maxpool = nn.MaxPool2d(2)
            inplanes = 256
            dilations = [1, 6, 12, 18]
            self.aspp1 = _ASPPModule(inplanes, 256, 1, padding=0, dilation=dilations[0])
            self.aspp2 = _ASPPModule(inplanes, 256, 3, padding=dilations[1], dilation=dilations[1])
            self.aspp3 = _ASPPModule(inplanes, 256, 3, padding=dilations[2], dilation=dilations[2])
            self.aspp4 = _ASPPModule(inplanes, 256, 3, padding=dilations[3], dilation=dilations[3])
            self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)),
                                                 nn.Conv2d(inplanes, 256, 1, stride=1, bias=False),
                                                 nn.BatchNorm2d(256), nn.ReLU())
            self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)
            self.conv2 = nn.Conv2d(2 * f, f, 1, bias=False)
            self.conv3 = nn.Conv2d(f, f, 3, 2, 1)

        def _emit(self, b, x):
            up1 = x
            low1 = self.residual_block_stride._emit(b, x)
            low2 = self.hourglass1._emit(b, low1) if self.n > 1 else low1
            low3 = self.residual_block._emit(b, low2)
            up2 = b.upsample2x_add(low3, None, mode="bilinear")
            return b.conv_cat(self.conv2, [up1, up2])

    class creatModel(HGModule):
        _is_model = True

        def __init__(self):
            super(creatModel, self).__init__()
            nFeats = g["nFeats"]
            self.conv1 = nn.Conv2d(3, 64, 7, 2, 3)
            self.relu = nn.ReLU()
            self.residual1 = ResidualBlock(64, 128, stride=2)
            self.residual2 = ResidualBlock(128, 128)
            self.residual3 = ResidualBlock(128, nFeats)
            self.hourglass1 = hourglass(4, nFeats)
            self.residual4 = ResidualBlock(nFeats, nFeats)
            self.lin = lin(nFeats, nFeats)
            self.conv2_0 = nn.Conv2d(nFeats, g["nOutChannels_0"], 1, 1, 0, bias=False)
            self.conv4_0 = nn.Conv2d(2 * nFeats + g["nOutChannels_0"], nFeats, 1, 1, 0)
            self.conv2_1 = nn.Conv2d(nFeats, g["nOutChannels_1"], 1, 1, 0, bias=False)
            self.conv4_1 = nn.Conv2d(2 * nFeats + g["nOutChannels_1"], nFeats, 1, 1, 0, bias=False)
            self.conv2_2 = nn.Conv2d(nFeats, g["nOutChannels_2"], 1, 1, 0, bias=False)

        def _config_key(self):
            return (g["nStack"],)

        def _emit(self, b, x):
            return _multihead_forward(self, b, x, g, cat_inter=True, with_pool=False, n_res4=1)

    for cls in (hourglass, creatModel, _ASPPModule):
        cls.__module__ = g.get("__name__", cls.__module__)
        cls.__qualname__ = cls.__name__
    return ResidualBlock, hourglass, lin, creatModel, _ASPPModule


def make_u_family(g):
    """hourglass_compare.py:405-638 (= creatModel_hourglass of performance_compare.py:335-427): the un-shared
    4-stage 'stacked hourglass' baseline on MPII-16 -- explicit 4-level hourglass with separate blocks, NEAREST
    up-sampling, stem conv + BN + ReLU, bias-free heads, inter = return(out) + inter + down_feature(ll)."""
    ResidualBlock = make_q4_block(g)

    class hourglass(HGModule):
        def __init__(self, f):
            super(hourglass, self).__init__()
            self.f = f
            self.downsample1 = nn.Sequential(nn.MaxPool2d(2, 2), ResidualBlock(f, f))
            self.downsample2 = nn.Sequential(nn.MaxPool2d(2, 2), ResidualBlock(f, f))
            self.downsample3 = nn.Sequential(nn.MaxPool2d(2, 2), ResidualBlock(f, f))
            self.downsample4 = nn.Sequential(nn.MaxPool2d(2, 2), ResidualBlock(f, f))
            self.residual1 = ResidualBlock(f, f)
            self.residual2 = ResidualBlock(f, f)
            self.residual3 = ResidualBlock(f, f)
            self.residual4 = ResidualBlock(f, f)
            self.residual5 = ResidualBlock(f, f)
            self.upsample1 = ResidualBlock(f, f)
            self.upsample2 = ResidualBlock(f, f)
            self.upsample3 = ResidualBlock(f, f)
            self.upsample4 = ResidualBlock(f, f)

        def _emit(self, b, x):
            with b.on_lane(4):
                up1 = self.residual1._emit(b, x)
            down1 = self.downsample1[1]._emit(b, b.maxpool2(x))
            with b.on_lane(3):
                up2 = self.residual2._emit(b, down1)
            down2 = self.downsample2[1]._emit(b, b.maxpool2(down1))
            with b.on_lane(2):
                up3 = self.residual3._emit(b, down2)
            down3 = self.downsample3[1]._emit(b, b.maxpool2(down2))
            with b.on_lane(1):
                up4 = self.residual4._emit(b, down3)
            down4 = self.downsample4[1]._emit(b, b.maxpool2(down3))
            out = self.residual5._emit(b, down4)
            out = self.upsample4._emit(b, out)
            out = b.upsample2x_add(out, up4, mode="nearest")
            out = self.upsample3._emit(b, out)
            out = b.upsample2x_add(out, up3, mode="nearest")
            out = self.upsample2._emit(b, out)
            out = b.upsample2x_add(out, up2, mode="nearest")
            out = self.upsample1._emit(b, out)
            out = b.upsample2x_add(out, up1, mode="nearest")
            return out

    class creatModel(HGModule):
        _is_model = True

        def __init__(self):
            super(creatModel, self).__init__()
            nFeats = g["nFeats"]
            self.preprocess1 = nn.Sequential(nn.Conv2d(3, 64, 7, 2, 3), nn.BatchNorm2d(64), nn.ReLU(),
                                             ResidualBlock(64, 128), nn.MaxPool2d(2, 2), ResidualBlock(128, 128),
                                             ResidualBlock(128, nFeats))
            for k in (1, 2, 3, 4):
                setattr(self, f"stage{k}", nn.Sequential(hourglass(nFeats), ResidualBlock(nFeats, nFeats),
                                                         nn.Conv2d(nFeats, nFeats, 1, 1, 0), nn.BatchNorm2d(nFeats),
                                                         nn.ReLU()))
                setattr(self, f"stage{k}_out", nn.Conv2d(nFeats, 16, 1, 1, 0, bias=False))
                if k < 4:
                    setattr(self, f"stage{k}_return", nn.Conv2d(16, nFeats, 1, 1, 0, bias=False))
                    setattr(self, f"stage{k}_down_feature", nn.Conv2d(nFeats, nFeats, 1, 1, 0, bias=False))

        def _emit(self, b, x):
            pre = self.preprocess1
            x = b.stem(pre[0], x, relu=False)
            x = b.bn_relu(pre[1], x)
            x = pre[3]._emit(b, x)
            x = b.maxpool2(x)
            x = pre[5]._emit(b, x)
            inter = pre[6]._emit(b, x)
            out = []
            for k in (1, 2, 3, 4):
                stage = getattr(self, f"stage{k}")
                ll = stage[0]._emit(b, inter)
                ll = stage[1]._emit(b, ll)
                ll = b.bn_relu(stage[3], b.conv(stage[2], ll))
                tmpOut = b.conv(getattr(self, f"stage{k}_out"), ll, head=True)
                out.insert(k - 1, tmpOut)
                if k < 4:
                    ll_ = b.conv(getattr(self, f"stage{k}_down_feature"), ll, residual=inter)
                    inter = b.conv(getattr(self, f"stage{k}_return"), tmpOut, residual=ll_)
            return out

    for cls in (hourglass, creatModel):
        cls.__module__ = g.get("__name__", cls.__module__)
        cls.__qualname__ = cls.__name__
    return ResidualBlock, hourglass, creatModel
