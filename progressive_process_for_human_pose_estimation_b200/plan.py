"""Execution plans: the host side of the hourglass hot path.

A reference model is a Python call tree of tiny nn.Modules (716 convolutions + 689 BatchNorms per 8-stack
forward, try_with_torch.py:275-298).  Driving that many kernels from autograd would leave the GPU idle, so the
drop-in modules (see _family_s.py) do not execute op by op.  On first call for a given input shape a module
*emits* its forward into a `Builder` (a small SSA graph of NHWC activations), the graph is *lowered* once to a
flat list of pre-marshalled C-ABI calls (forward and hand-derived backward) on statically allocated buffers,
and both lists are captured into CUDA graphs.  A training step is then: one graph launch for the forward, the
stock loss on the returned heatmaps, one graph launch for the backward, the stock optimizer.

All arithmetic happens in libhg_sm100a.so (include/hg_sm100a.h); nothing here computes on tensors except
buffer bookkeeping (zero-fill, input copy, output clone).
"""
import ctypes as C
import os

import torch

from . import _lib as L

_USE_GRAPHS = os.environ.get("HG_CUDA_GRAPHS", "1") != "0"
_USE_LANES = os.environ.get("HG_STREAM_LANES", "1") != "0"
_WGRAD_LANES = int(os.environ.get("HG_WGRAD_LANES", "4"))
# 0: BatchNorm kernels stay separate; 1: only the data-gradient epilogue is fused (ReLU mask + BN-backward sums);
# 2: additionally the forward / weight-gradient convolutions apply BN+ReLU to their operand tiles (no activation in HBM)
# 3: like 2, but only where the consumer is a 1x1 convolution (bn1 -> conv1, bn3 -> conv3 of a residual block)
# 4: 3 + every consumer whose grid is a single wave (the 4x4 .. 16x16 levels are latency chains: a separate BatchNorm
#    kernel costs a launch there, the shared-memory rewrite nothing);  5: only the single-wave consumers
_FOLD_BN = int(os.environ.get("HG_FOLD_BN", "1"))
# inference (eval mode, no gradients): a BatchNorm(+ReLU) whose producer is a tensor-core convolution runs in that
# convolution's epilogue (hg_conv_fprop_bnout)
_FUSE_EVAL_BN = os.environ.get("HG_FUSE_EVAL_BN", "1") != "0"
# Stream priority of the main lane in the FORWARD graph: it carries the latency-bound low-resolution chain, and when a
# skip-branch lane's big kernel holds every SM the CTA scheduler must hand freed slots to the main lane first
# (measured: forward 14.3 -> 13.1 ms).
_FWD_MAIN_PRIORITY = int(os.environ.get("HG_FWD_MAIN_PRIORITY", "-3"))
# Backward graph: the main lane (dgrad / BatchNorm chain with its latency-bound low-resolution stretches) above the
# skip-branch and wgrad lanes, which fill the SMs it leaves idle; several wgrad lanes keep enough of that deferred work
# in flight (4 lanes: backward 27.45 -> 26.9 ms; skip lanes at normal priority too: 850.5 -> 856.8 images/s).
_BWD_PRIO = [int(v) for v in os.environ.get("HG_BWD_PRIORITY", "-3,0,0").split(",")]  # main, skip lanes, wgrad lanes
# 1: (round 1) the packed-gradient unpack of every shared weight runs on the main lane behind a barrier over all lanes
_UNPACK_BARRIER = os.environ.get("HG_UNPACK_BARRIER", "0") == "1"
# 1: the per-step weight repacking runs on a side lane under the stem convolution (which needs no packed weight)
_PACK_SIDE_LANE = os.environ.get("HG_PACK_SIDE_LANE", "1") == "1"
# 1: max-pool backward lowered after the other gradient contributions of its input (it adds into their sum: the 32
# hg_add launches of a step disappear, 134 MB less traffic each at 64x64).  Measured: backward phase 23.3 -> 22.8 ms but
# the steady-state step 35.25 -> 35.5 ms (the pool backward then sits at the END of the level's gradient chain instead
# of running early next to it), so it stays off.
_DEFER_POOL_BWD = os.environ.get("HG_DEFER_POOL_BWD", "0") == "1"


class Val:
    """An NHWC activation [N,H,W,Cp] (channels padded to 64) and, during lowering, its gradient."""

    __slots__ = ("N", "H", "W", "C", "buf", "grad", "grad_owned", "requires_grad", "needs_stats", "stats",
                 "producer", "consumers", "name")

    def __init__(self, N, H, W, C, requires_grad, name=""):
        self.N, self.H, self.W, self.C = N, H, W, C
        self.buf = None
        self.grad = None
        self.grad_owned = False
        self.requires_grad = requires_grad
        self.needs_stats = False
        self.stats = None
        self.producer = None
        self.consumers = []
        self.name = name

    @property
    def Cp(self):
        return L.pad64(self.C)

    @property
    def M(self):
        return self.N * self.H * self.W

    def numel_padded(self):
        return self.M * self.Cp


class Op:
    lane = 0

    def __init__(self, kind, ins, out, **attrs):
        self.kind = kind
        self.ins = ins
        self.out = out
        self.attrs = attrs
        for v in ins:
            if v is not None:
                v.consumers.append(self)
        if out is not None:
            out.producer = self


class Builder:
    """Collects the ops a module tree emits.  Methods mirror the nn calls of the reference forward."""

    def __init__(self, training, train_params):
        self.training = training
        self.train_params = train_params  # parameters require grad -> activations feeding them too
        self.ops = _LaneList(self)
        self.inputs = []    # (Val, kind)
        self.outputs = []   # (Val, real channels)
        self.cur_lane = 0   # stream lane new ops are assigned to (0 = main)
        self.num_lanes = 1

    def on_lane(self, lane):
        """Context manager: ops emitted inside run on stream lane `lane` (independent branches of the network --
        the skip branch of an hourglass level vs. its low-resolution path -- execute concurrently)."""
        return _LaneCtx(self, lane)

    # -- graph inputs -------------------------------------------------------------------------------
    def input_image(self, N, H, W):
        """fp32 NCHW image batch consumed directly by the stem kernel."""
        v = Val(N, H, W, 3, False, "image")
        self.inputs.append((v, "image"))
        return v

    def input_nchw(self, N, C, H, W, requires_grad):
        v = Val(N, H, W, C, requires_grad, "input")
        self.inputs.append((v, "nchw"))
        return v

    def _rg(self, *vals):
        return self.train_params or any(v.requires_grad for v in vals if v is not None)

    # -- ops ----------------------------------------------------------------------------------------
    def stem(self, conv, x, relu=True):
        """7x7 stride-2 stem on the fp32 NCHW image; relu=False when a BatchNorm follows (hourglass_compare.py:549)."""
        assert conv.kernel_size == (7, 7) and conv.stride == (2, 2) and conv.padding == (3, 3)
        out = Val(x.N, x.H // 2, x.W // 2, conv.out_channels, self.train_params, "stem")
        self.ops.append(Op("stem", [x], out, conv=conv, relu=relu))
        return out

    def conv(self, conv, x, residual=None, head=False, cin_off=0, use_bias=True, mix=None):
        """nn.Conv2d (+ fused residual add).  head=True also produces the fp32 NCHW tensor the module returns.
        cin_off: x feeds the input-channel slice [cin_off, cin_off + x.C) of the weight (see conv_cat).
        mix: [Cout', Cout] matrix T applied to the output channels (folded into the weights: T W, T b); the op then
        produces Cout' channels (square for the in-place limb mix, [36, 17] for the gather-add limb maps)."""
        k, s, p, d = conv.kernel_size, conv.stride, conv.padding, conv.dilation
        assert k[0] == k[1] and s[0] == s[1] and p[0] == p[1] and d[0] == d[1]
        assert cin_off + x.C <= conv.in_channels and (cin_off > 0 or x.C <= conv.in_channels), (conv.in_channels, x.C)
        Ho = (x.H + 2 * p[0] - d[0] * (k[0] - 1) - 1) // s[0] + 1
        Wo = (x.W + 2 * p[0] - d[0] * (k[0] - 1) - 1) // s[0] + 1
        cout = conv.out_channels if mix is None else int(mix.shape[0])
        assert mix is None or int(mix.shape[1]) == conv.out_channels
        out = Val(x.N, Ho, Wo, cout, self._rg(x, residual), "conv")
        self.ops.append(Op("conv", [x, residual], out, conv=conv, head=head, cin_off=cin_off,
                           use_bias=use_bias and conv.bias is not None, mix=mix, cout=cout))
        if head:
            self.outputs.append((out, cout))
        return out

    def conv_cat(self, conv, xs, head=False):
        """conv(torch.cat(xs, 1)) without materialising the concatenation: one chained convolution per input over
        the matching slice of the weight, each adding to the previous partial sum through the residual epilogue."""
        assert sum(v.C for v in xs) == conv.in_channels, ([v.C for v in xs], conv.in_channels)
        y, off = None, 0
        for i, v in enumerate(xs):
            y = self.conv(conv, v, residual=y, head=head and i == len(xs) - 1, cin_off=off, use_bias=(i == 0))
            off += v.C
        return y

    def bn_relu(self, bn, x, relu=True):
        assert bn.num_features == x.C
        # per module, like nn.BatchNorm2d itself: a sub-module put in eval() under a training root uses (and keeps) its
        # running statistics
        if bn.training or not bn.track_running_stats:
            x.needs_stats = True
        out = Val(x.N, x.H, x.W, x.C, self._rg(x), "bn")
        self.ops.append(Op("bn", [x], out, bn=bn, relu=relu))
        return out

    def maxpool2(self, x):
        out = Val(x.N, x.H // 2, x.W // 2, x.C, x.requires_grad, "pool")
        self.ops.append(Op("pool", [x], out))
        return out

    def upsample2x_add(self, low, skip, mode="bilinear"):
        out = Val(low.N, low.H * 2, low.W * 2, low.C, self._rg(low, skip), "up")
        self.ops.append(Op("up", [low, skip], out, mode=0 if mode == "bilinear" else 1))
        return out

    def global_avg_pool(self, x):
        """nn.AdaptiveAvgPool2d((1, 1)) (train.py:476)."""
        out = Val(x.N, 1, 1, x.C, self._rg(x), "gap")
        self.ops.append(Op("gap", [x], out))
        return out

    def broadcast_to(self, y, H, W):
        """F.interpolate of a 1x1 map to HxW (bilinear, align_corners=True: a constant map; train.py:489)."""
        assert y.H == 1 and y.W == 1
        out = Val(y.N, H, W, y.C, self._rg(y), "bcast")
        self.ops.append(Op("bcast", [y], out))
        return out

    def cat(self, xs):
        """torch.cat(xs, dim=1) as a real tensor (its consumer is a BatchNorm: train.py:528-538,570-583).  Every
        input but the last needs a channel count that is a multiple of 8."""
        assert all(v.C % 8 == 0 for v in xs[:-1]), [v.C for v in xs]
        assert all((v.N, v.H, v.W) == (xs[0].N, xs[0].H, xs[0].W) for v in xs)
        out = Val(xs[0].N, xs[0].H, xs[0].W, sum(v.C for v in xs), self._rg(*xs), "cat")
        self.ops.append(Op("cat", list(xs), out))
        return out

    def add(self, a, b):
        out = Val(a.N, a.H, a.W, a.C, self._rg(a, b), "add")
        self.ops.append(Op("add", [a, b], out))
        return out

    def output(self, v):
        """Return `v` to the caller as an fp32 NCHW tensor."""
        self.outputs.append((v, v.C))
        self.ops.append(Op("export", [v], None))
        return v


class _LaneList(list):
    """Op list that stamps every appended op with the builder's current lane."""

    def __init__(self, builder):
        super().__init__()
        self._b = builder

    def append(self, op):
        op.lane = self._b.cur_lane
        super().append(op)


class _LaneCtx:
    def __init__(self, b, lane):
        self.b, self.lane = b, lane

    def __enter__(self):
        self.prev = self.b.cur_lane
        self.b.cur_lane = self.lane
        self.b.num_lanes = max(self.b.num_lanes, self.lane + 1)

    def __exit__(self, *exc):
        self.b.cur_lane = self.prev
        return False


# positions (in the ctypes argument tuple) of the tensors each entry point WRITES; every other tracked pointer
# argument is a read.  Parameter / weight-gradient pointers are not tracked (read-only or commutative atomics).
_WRITES = {
    "hg_nchw_f32_to_nhwc": (7,), "hg_nhwc_to_nchw_f32": (6,), "hg_stem_fwd": (8,), "hg_conv_fprop_ex": (5, 6, 7),
    "hg_conv_dgrad": (4,), "hg_conv_fprop_bnout": (5, 6), "hg_bn_stats": (2,), "hg_bn_apply": (7,), "hg_bn_bwd_apply": (10,),
    "hg_bn_bwd_reduce": (8,), "hg_conv_fprop_bn": (6, 7, 8), "hg_conv_dgrad_bn": (5, 6),
    "hg_maxpool2_fwd": (6, 7), "hg_maxpool2_bwd": (8,), "hg_upsample2x_add_fwd": (8, 9), "hg_upsample2x_bwd": (8,),
    "hg_add": (3,), "hg_spatial_mean": (8,), "hg_spatial_broadcast": (8,),
    "hg_channel_copy": (5,),
}


class _Call:
    """One pre-marshalled C-ABI call."""

    __slots__ = ("fn", "args", "name", "keep", "writes", "tag", "lane", "deps", "event", "barrier")

    def __init__(self, name, args, keep=()):
        self.fn = getattr(L.load(), name)
        self.args = args
        self.name = name
        self.keep = keep  # python objects whose memory the ctypes args point to
        self.writes = ()  # indices of parameters whose gradient slot this call writes
        self.tag = ""     # human-readable shape tag for the per-kernel profile (bench.py)
        self.lane = 0     # stream lane
        self.deps = ()    # calls on OTHER lanes that must complete first
        self.event = None  # recorded after this call when another lane depends on it
        self.barrier = False  # wait for every lane before this call


def _unique(seq):
    seen, out = set(), []
    for s in seq:
        if id(s) not in seen:
            seen.add(id(s))
            out.append(s)
    return out


class Plan:
    def __init__(self, builder, params, device, compute_dtype):
        """params: ordered list of (name, nn.Parameter) of the root module (the autograd inputs)."""
        self.b = builder
        self.device = device
        self.dt = compute_dtype
        self.hdt = L.hg_dtype(compute_dtype)
        self.params = params
        self.param_index = {id(p): i for i, (_, p) in enumerate(params)}
        self.training = builder.training
        self.need_bwd = builder.train_params or any(v.requires_grad for v, _ in builder.inputs)
        self.stream = C.c_void_p(0)
        self._keep = []
        self.fwd_calls = []
        self.bwd_calls = []
        self.fwd_graph = None
        self.bwd_graphs = {}
        self.n_fwd_runs = 0
        self._pending_writes = []
        self._tracked = set()       # data_ptrs of activation / gradient / statistics buffers (dependency tracking)
        self._last_writer = {}      # data_ptr -> _Call that last wrote it (per call list)
        self._cur_lane = 0
        self.num_lanes = builder.num_lanes if _USE_LANES else 1
        self.wgrad_lanes = []
        self._wgrad_rr = 0
        if _USE_LANES and _WGRAD_LANES > 0:
            self.wgrad_lanes = list(range(self.num_lanes, self.num_lanes + _WGRAD_LANES))
            self.num_lanes += _WGRAD_LANES
        self.lane_streams = {}     # direction -> side-lane streams (the two graphs use different priorities)
        self.profile_records = None  # list while an instrumented (eager, event-timed) step is being recorded
        self.reducer = None        # set by parallel.DataParallel: all-reduces ranges of grad_arena
        self.bwd_segments = None   # [(first_call, end_call, [(lo, hi) element ranges of grad_arena ready after it])]
        self._lower()

    # ------------------------------------------------------------------------------------------------
    # buffer helpers
    # ------------------------------------------------------------------------------------------------
    def _act(self, v):
        t = torch.zeros(v.N, v.H, v.W, v.Cp, device=self.device, dtype=self.dt)
        self._tracked.add(t.data_ptr())
        return t

    def _p32(self, p):
        """fp32 view of a parameter / buffer (shadow copy when the module was cast to half)."""
        if p.dtype == torch.float32:
            return p.data
        key = id(p)
        if key not in self._shadow:
            self._shadow[key] = (p, torch.empty_like(p.data, dtype=torch.float32))
        return self._shadow[key][1]

    def _rstat(self, buf):
        """fp32 tensor behind a BatchNorm running-statistics buffer: the buffer itself, or -- when the module was
        cast with .half() as the reference's test paths do (try_with_torch.py:374-377) -- an fp32 shadow that is
        refreshed before and written back after every forward."""
        if buf is None:
            return None
        if buf.dtype == torch.float32:
            return buf
        key = id(buf)
        if key not in self._shadow_buf:
            self._shadow_buf[key] = (buf, torch.empty_like(buf, dtype=torch.float32))
        return self._shadow_buf[key][1]

    def _gslot(self, p):
        """fp32 gradient slot of parameter p inside the flat gradient arena (None if p is frozen)."""
        i = self.param_index.get(id(p))
        if i is None or not p.requires_grad:
            return None
        self.param_used[i] = True
        self._pending_writes.append(i)
        return self.grad_views[i]

    def _gslot_peek(self, p):
        """True when parameter p receives a gradient (without registering a write)."""
        i = self.param_index.get(id(p))
        return i is not None and p.requires_grad

    def _emit(self, lst, name, *args, keep=(), reads=()):
        """reads: tensors the call reads through a struct (HgBnFold) rather than through a pointer argument."""
        c = _Call(name, args, keep)
        if lst is self.bwd_calls and self._pending_writes:
            c.writes = tuple(self._pending_writes)
            self._pending_writes = []
        c.lane = self._cur_lane if self.num_lanes > 1 else 0
        # cross-lane dependencies from the tracked buffers this call reads / writes
        wpos = _WRITES.get(name, ())
        deps = []
        for i, a in enumerate(args):
            if isinstance(a, C.c_void_p) and a.value in self._tracked:
                lw = self._last_writer.get(a.value)
                if lw is not None and lw.lane != c.lane and lw not in deps:
                    deps.append(lw)
                if i in wpos:
                    self._last_writer[a.value] = c
        for t in reads:
            lw = self._last_writer.get(t.data_ptr()) if t is not None else None
            if lw is not None and lw.lane != c.lane and lw not in deps:
                deps.append(lw)
        pd = getattr(self, "_pack_done", None)
        if (lst is self.fwd_calls and pd is not None and name.startswith("hg_conv_") and c.lane != pd.lane
                and c.lane not in self._pack_synced):
            self._pack_synced.add(c.lane)   # later calls of the lane follow in stream order
            if pd not in deps:
                deps.append(pd)
        c.deps = tuple(deps)
        lst.append(c)
        return c

    # ------------------------------------------------------------------------------------------------
    # gradient bookkeeping (see module docstring of _family_s for the conventions)
    # ------------------------------------------------------------------------------------------------
    def _grad_target(self, v):
        """(addend, out) buffers for a kernel that ADDS a computed contribution to v.grad."""
        if v.grad is None:
            v.grad = self._act(v)
            v.grad_owned = True
            return None, v.grad
        if v.grad_owned:
            return v.grad, v.grad
        old = v.grad
        v.grad = self._act(v)
        v.grad_owned = True
        return old, v.grad

    def _grad_passthrough(self, v, g):
        """v.grad += g where g is an existing buffer (no kernel when it is the first contribution)."""
        if not v.requires_grad:
            return
        if v.grad is None:
            v.grad = g
            v.grad_owned = False
            return
        addend, out = self._grad_target(v)
        self._emit(self.bwd_calls, "hg_add", self.hdt, L.ptr(addend), L.ptr(g), L.ptr(out),
                   C.c_longlong(v.numel_padded()), self.stream)

    # ------------------------------------------------------------------------------------------------
    def _lower(self):
        b, dev = self.b, self.device
        self._shadow = {}
        self._shadow_buf = {}
        ops = b.ops
        # ---- parameters: flat gradient arena in named_parameters() order --------------------------
        # every slot starts on a 16-byte boundary: the wgrad kernel reduces straight into the slots of un-padded
        # 1x1 weights with red.global.add.v4.f32
        sizes = [p.numel() for _, p in self.params]
        self.grad_offsets, off = [], 0
        for n in sizes:
            self.grad_offsets.append(off)
            off += (n + 3) // 4 * 4
        self.grad_arena = torch.zeros(max(4, off), device=dev, dtype=torch.float32)
        self.grad_views = [self.grad_arena[o:o + n].view(p.shape)
                           for (_, p), n, o in zip(self.params, sizes, self.grad_offsets)]
        self.param_used = [False] * len(self.params)

        # ---- convolutions: packed operands, packed gradient accumulators -------------------------
        # one entry per (conv module, input-channel slice): a shared weight has ONE entry for all its call sites
        conv_ops = [op for op in ops if op.kind == "conv"]
        self.conv_info = {}
        self.conv_keys = []
        packed_total = 0
        for op in conv_ops:
            cv, x = op.attrs["conv"], op.ins[0]
            key = (id(cv), op.attrs["cin_off"])
            op.attrs["key"] = key
            if key in self.conv_info:
                assert self.conv_info[key]["cin_seg"] == x.C
                continue
            k = cv.kernel_size[0]
            mix = op.attrs["mix"]
            cout = op.attrs["cout"]   # != cv.out_channels under a rectangular channel mix
            cin_p, cout_p = L.pad64(x.C), L.pad64(cout)
            whole = (op.attrs["cin_off"] == 0 and x.C == cv.in_channels)
            direct = (k == 1 and whole and mix is None and cin_p == x.C and cout_p == cv.out_channels)
            info = dict(conv=cv, cin_off=op.attrs["cin_off"], cin_seg=x.C, mix=None, cout=cout,
                        wf=torch.zeros(k * k, cout_p, cin_p, device=dev, dtype=self.dt),
                        wd=torch.zeros(k * k, cin_p, cout_p, device=dev, dtype=self.dt),
                        bias=torch.zeros(cout_p, device=dev, dtype=torch.float32), direct=direct,
                        gsize=k * k * cout_p * cin_p, goff=packed_total)
            if not direct:
                packed_total += info["gsize"]
            if mix is not None:
                assert whole, "channel mixing is only supported on un-sliced convolutions"
                info["mix"] = mix.to(device=dev, dtype=torch.float32).contiguous()
                info["w_eff"] = torch.zeros(cout, cv.in_channels * k * k, device=dev)
                info["dweff_off"] = packed_total
                packed_total += cout * cv.in_channels * k * k
                info["dbeff_off"] = packed_total
                packed_total += cout_p
                info["bias_copy"] = False
            self.conv_info[key] = info
            self.conv_keys.append(key)
        convs = self.conv_keys
        self.packed_arena = torch.zeros(max(1, packed_total), device=dev, dtype=torch.float32)

        # ---- activations and BatchNorm statistics ---------------------------------------------
        stat_vals = []
        for v, kind in b.inputs:
            if kind == "image":
                v.buf = torch.zeros(v.N, 3, v.H, v.W, device=dev, dtype=torch.float32)
            else:
                v.buf = self._act(v)
                self.in_nchw = torch.zeros(v.N, v.C, v.H, v.W, device=dev, dtype=torch.float32)
        # ---- BatchNorm folding: BN(+ReLU) whose only consumer is a tensor-core convolution is never materialised;
        # the convolution transforms its operand tiles on the fly (hg_conv_fprop_bn / _wgrad_bn / _dgrad_bn)
        self.n_folded = self.n_masked = 0
        if _FOLD_BN and self.dt == torch.bfloat16:
            lib = L.load()
            for op in ops:
                if op.kind != "bn" or op.out.needs_stats or len(op.out.consumers) != 1:
                    continue
                cons = op.out.consumers[0]
                if cons.kind != "conv" or cons.ins[0] is not op.out or cons.ins[1] is op.out:
                    continue
                if not lib.hg_conv_fold_eligible(C.byref(self._conv_desc(cons.attrs["conv"], op.out))):
                    continue
                # activation not materialised: every consumer (2), or only 1x1 consumers (3: one shared-memory rewrite per
                # operand element instead of nine for a 3x3)
                k1 = cons.attrs["conv"].kernel_size[0] == 1
                one_wave = (op.out.M + 127) // 128 * max(1, L.pad64(cons.attrs["conv"].out_channels) // 128) <= 148
                op.attrs["folded"] = (_FOLD_BN == 2 or (_FOLD_BN == 3 and k1) or (_FOLD_BN == 4 and (k1 or one_wave))
                                      or (_FOLD_BN == 5 and one_wave))
                op.attrs["masked"] = True            # consumer's dgrad epilogue does the reduce
                cons.attrs["fold"] = op
                self.n_masked += 1
                self.n_folded += 1 if op.attrs["folded"] else 0
        # ---- inference: conv -> eval-mode BN(+ReLU) in ONE kernel (hg_conv_fprop_bnout): the BatchNorm is a per-channel
        # affine with known coefficients, applied in the convolution's epilogue; the raw conv output never exists
        self.n_bn_out = 0
        if _FUSE_EVAL_BN and self.dt == torch.bfloat16 and not self.need_bwd:
            lib = L.load()
            for op in ops:
                if op.kind != "conv" or op.ins[1] is not None or op.attrs["head"] or op.attrs["mix"] is not None:
                    continue
                out = op.out
                if out.needs_stats or len(out.consumers) != 1 or any(out is v for v, _ in b.outputs):
                    continue
                cons = out.consumers[0]
                if cons.kind != "bn" or cons.attrs.get("folded") or not cons.attrs["bn"].track_running_stats:
                    continue
                if cons.attrs["bn"].training:
                    continue
                if op.attrs.get("fold") is not None and op.attrs["fold"].attrs.get("folded"):
                    continue
                if not lib.hg_conv_tc_eligible(C.byref(self._conv_desc(op.attrs["conv"], op.ins[0]))):
                    continue
                op.attrs["bn_out"] = cons
                cons.attrs["in_producer"] = True
                self.n_bn_out += 1
        for op in ops:
            if op.out is not None and not op.attrs.get("folded") and op.attrs.get("bn_out") is None:
                op.out.buf = self._act(op.out)
        all_vals = [v for v, _ in b.inputs] + [op.out for op in ops if op.out is not None]
        for v in all_vals:
            if v.needs_stats:
                stat_vals.append(v)
        # statistics slots: {S1, S2, pivot}[3*Cp] per tensor (shifted sums, include/hg_sm100a.h); the pivot is the running
        # mean of the BatchNorm that consumes the tensor, written into the slot at the start of every forward by
        # hg_bn_prepare_stats (which also zeroes S1 / S2: no separate memset of the arena)
        n_stats = sum(3 * v.Cp for v in stat_vals)
        self.stats_arena = torch.zeros(max(1, n_stats), device=dev, dtype=torch.float32)
        off = 0
        slots = []
        for v in stat_vals:
            v.stats = self.stats_arena[off:off + 3 * v.Cp]
            self._tracked.add(v.stats.data_ptr())
            off += 3 * v.Cp
            pivot = None
            for cons in v.consumers:
                if cons.kind == "bn" and cons.attrs["bn"].running_mean is not None:
                    pivot = self._rstat(cons.attrs["bn"].running_mean)
                    break
            slots.append(L.HgBnStatsSlot(v.stats.data_ptr(), pivot.data_ptr() if pivot is not None else None, v.C, v.Cp))
        self.stats_slots = None
        if slots:
            arr = (L.HgBnStatsSlot * len(slots))(*slots)
            self.stats_slots = (torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev), len(slots))
        bn_ops = [op for op in ops if op.kind == "bn"]
        self.red_arena = torch.zeros(max(1, sum(2 * op.ins[0].Cp for op in bn_ops)), device=dev, dtype=torch.float32)
        off = 0
        for op in bn_ops:
            op.attrs["red"] = self.red_arena[off:off + 2 * op.ins[0].Cp]
            self._tracked.add(op.attrs["red"].data_ptr())
            off += 2 * op.ins[0].Cp

        # ---- outputs -------------------------------------------------------------------------------
        # the tensors the module returns: views of ONE arena, so that handing the caller its own copies is a single
        # device-to-device copy per forward (not one per stack), and the incoming gradients land in a second arena
        shapes = [(v.N, c, v.H, v.W) for v, c in b.outputs]
        numels = [(n * c * h * w + 3) // 4 * 4 for n, c, h, w in shapes]   # 16-byte aligned slots
        self.out_arena = torch.zeros(max(4, sum(numels)), device=dev, dtype=torch.float32)
        self.out_static, off = [], 0
        for shp, n in zip(shapes, numels):
            self.out_static.append(self.out_arena[off:off + shp[0] * shp[1] * shp[2] * shp[3]].view(shp))
            off += n
        self._out_slices = [(o, shp) for o, shp in zip([sum(numels[:i]) for i in range(len(numels))], shapes)]
        self.gout_static = []
        if self.need_bwd:
            self.gout_arena = torch.zeros_like(self.out_arena)
            for o, shp in self._out_slices:
                self.gout_static.append(self.gout_arena[o:o + shp[0] * shp[1] * shp[2] * shp[3]].view(shp))
        self.out_index = {id(v): i for i, (v, _) in enumerate(b.outputs)}
        for t in self.out_static + self.gout_static:
            self._tracked.add(t.data_ptr())

        self._lower_forward(convs)
        if self.need_bwd:
            self._lower_backward(convs)
        for c in self.fwd_calls + self.bwd_calls:   # shape tags of the BatchNorm calls (per-kernel profile)
            if c.name in ("hg_bn_apply", "hg_bn_bwd_apply", "hg_bn_bwd_reduce", "hg_bn_stats") and not c.tag:
                d = c.args[0]._obj
                c.tag = f"C{d.C} M{d.M}" + (" +addend" if c.name == "hg_bn_bwd_apply" and c.args[9] is not None else "")

    # ------------------------------------------------------------------------------------------------
    def _conv_desc(self, cv, x, cout=None):
        d = L.HgConvDesc(x.N, x.H, x.W, x.C, cv.out_channels if cout is None else cout, cv.kernel_size[0], cv.kernel_size[1],
                         cv.stride[0], cv.padding[0], cv.dilation[0], self.hdt)
        self._keep.append(d)
        return d

    @staticmethod
    def _conv_tag(cv, x):
        return f"{cv.in_channels}->{cv.out_channels} k{cv.kernel_size[0]} @{x.H}x{x.W}"

    def _bn_desc(self, bn, x, relu):
        use_running = 0 if (bn.training or not bn.track_running_stats) else 1
        d = L.HgBnDesc(x.M, x.C, self.hdt, float(bn.eps), 1 if relu else 0, use_running)
        self._keep.append(d)
        return d

    def _bn_fold(self, bnop):
        """HgBnFold of a folded BatchNorm op + the tensors a call using it reads through the struct."""
        bn, x = bnop.attrs["bn"], bnop.ins[0]
        use_running = 0 if (bn.training or not bn.track_running_stats) else 1
        f = L.HgBnFold(x.stats.data_ptr() if x.stats is not None else None, self._p32(bn.weight).data_ptr(),
                       self._p32(bn.bias).data_ptr(),
                       self._rstat(bn.running_mean).data_ptr() if bn.running_mean is not None else None,
                       self._rstat(bn.running_var).data_ptr() if bn.running_var is not None else None, float(bn.eps),
                       1 if bnop.attrs["relu"] else 0, use_running, 0)
        self._keep.append(f)
        return f, (x.stats,)

    def _lower_forward(self, convs):
        f, st = self.fwd_calls, self.stream
        if self.stats_slots is not None:
            self._emit(f, "hg_bn_prepare_stats", L.ptr(self.stats_slots[0]), self.stats_slots[1], st)
        # weights -> GEMM operand layouts (the optimizer changed them since the last step).  ~30 small launches: on a
        # side lane they run under the stem convolution, which reads the fp32 OIHW weight itself; the first convolution
        # of every lane waits for the last of them (`_pack_done`)
        self._pack_done, self._pack_synced = None, set()
        if _PACK_SIDE_LANE and self.wgrad_lanes and self.num_lanes > 1:
            self._cur_lane = self.wgrad_lanes[0]
        for key in convs:
            info = self.conv_info[key]
            cv = info["conv"]
            k = cv.kernel_size[0]
            d = L.HgConvDesc(1, 1, 1, info["cin_seg"], info["cout"], k, k, 1, 0, 1, self.hdt)
            self._keep.append(d)
            src = self._p32(cv.weight)
            if info["mix"] is not None:  # W_eff = T W, b_eff = T b
                self._emit(f, "hg_mix_rows_rect", L.ptr(info["mix"]), L.ptr(src), L.ptr(info["w_eff"]), info["cout"],
                           cv.out_channels, cv.in_channels * k * k, 0, 0, st)
                src = info["w_eff"]
                if cv.bias is not None:
                    self._emit(f, "hg_mix_rows_rect", L.ptr(info["mix"]), L.ptr(self._p32(cv.bias)), L.ptr(info["bias"]),
                               info["cout"], cv.out_channels, 1, 0, 0, st)
            self._pack_done = self._emit(f, "hg_pack_conv_weight_slice", C.byref(d), L.ptr(src), cv.in_channels,
                                         info["cin_off"], L.ptr(info["wf"]), L.ptr(info["wd"]) if self.need_bwd else None,
                                         st)
        self._cur_lane = 0
        for v, kind in self.b.inputs:
            if kind == "nchw":
                self._emit(f, "hg_nchw_f32_to_nhwc", self.hdt, L.ptr(self.in_nchw), None, v.N, v.C, v.H, v.W,
                           L.ptr(v.buf), st)
                if v.needs_stats:
                    self._stats_call(f, v)
        running = {}  # bn module -> list of (stats, count) in call order
        for op in self.b.ops:
            k = op.kind
            self._cur_lane = op.lane
            if k == "stem":
                cv, x, out = op.attrs["conv"], op.ins[0], op.out
                self._emit(f, "hg_stem_fwd", self.hdt, L.ptr(x.buf), L.ptr(self._p32(cv.weight)),
                           L.ptr(self._p32(cv.bias)) if cv.bias is not None else None, x.N, x.H, x.W,
                           1 if op.attrs["relu"] else 0, L.ptr(out.buf), st)
                if out.needs_stats:
                    self._stats_call(f, out)
            elif k == "conv":
                cv, x, res, out = op.attrs["conv"], op.ins[0], op.ins[1], op.out
                info = self.conv_info[op.attrs["key"]]
                d = self._conv_desc(cv, x, info["cout"])
                nchw = self.out_static[self.out_index[id(out)]] if op.attrs["head"] else None
                bias = self._bias_ptr(cv, info) if op.attrs["use_bias"] else None
                if op.attrs.get("bn_out") is not None:
                    bnop = op.attrs["bn_out"]
                    fold, _ = self._bn_fold(bnop)
                    self._emit(f, "hg_conv_fprop_bnout", C.byref(d), C.byref(fold), L.ptr(x.buf), L.ptr(info["wf"]),
                               bias, L.ptr(bnop.out.buf), None, st).tag = self._conv_tag(cv, x) + " +bn_out"
                elif op.attrs.get("fold") is not None and op.attrs["fold"].attrs["folded"]:
                    fold, reads = self._bn_fold(op.attrs["fold"])
                    self._emit(f, "hg_conv_fprop_bn", C.byref(d), C.byref(fold), L.ptr(op.attrs["fold"].ins[0].buf),
                               L.ptr(info["wf"]), bias, L.ptr(res.buf) if res else None, L.ptr(out.buf),
                               L.ptr(out.stats) if out.needs_stats else None, L.ptr(nchw), st,
                               reads=reads).tag = self._conv_tag(cv, x) + " +bn"
                else:
                    self._emit(f, "hg_conv_fprop_ex", C.byref(d), L.ptr(x.buf), L.ptr(info["wf"]), bias,
                               L.ptr(res.buf) if res else None,
                               L.ptr(out.buf), L.ptr(out.stats) if out.needs_stats else None, L.ptr(nchw),
                               st).tag = self._conv_tag(cv, x) + (" +res" if res else "")
            elif k == "bn":
                bn, x, out = op.attrs["bn"], op.ins[0], op.out
                d = self._bn_desc(bn, x, op.attrs["relu"])
                if not op.attrs.get("folded") and not op.attrs.get("in_producer"):
                    self._emit(f, "hg_bn_apply", C.byref(d), L.ptr(x.buf),
                               L.ptr(x.stats) if x.stats is not None else None,
                               L.ptr(self._p32(bn.weight)), L.ptr(self._p32(bn.bias)),
                               L.ptr(self._rstat(bn.running_mean)),
                               L.ptr(self._rstat(bn.running_var)), L.ptr(out.buf), st)
                if bn.training and bn.track_running_stats:
                    running.setdefault(id(bn), (bn, []))[1].append((x.stats, float(x.M)))
                if out.needs_stats:  # BN output feeding another BN directly (hourglass_compare.py:549-553)
                    self._stats_call(f, out)
            elif k == "pool":
                x, out = op.ins[0], op.out
                self._emit(f, "hg_maxpool2_fwd", self.hdt, L.ptr(x.buf), x.N, x.H, x.W, x.C, L.ptr(out.buf),
                           L.ptr(out.stats) if out.needs_stats else None, st)
            elif k == "up":
                low, skip, out = op.ins[0], op.ins[1], op.out
                self._emit(f, "hg_upsample2x_add_fwd", self.hdt, op.attrs["mode"], L.ptr(low.buf),
                           L.ptr(skip.buf) if skip else None, low.N, low.H, low.W, low.C, L.ptr(out.buf),
                           L.ptr(out.stats) if out.needs_stats else None, st)
            elif k == "add":
                a, bb, out = op.ins[0], op.ins[1], op.out
                self._emit(f, "hg_add", self.hdt, L.ptr(a.buf), L.ptr(bb.buf), L.ptr(out.buf),
                           C.c_longlong(out.numel_padded()), st)
                if out.needs_stats:
                    self._stats_call(f, out)
            elif k == "cat":
                out, off = op.out, 0
                for v in op.ins:
                    self._emit(f, "hg_channel_copy", self.hdt, L.ptr(v.buf), v.C, 0, None, L.ptr(out.buf), out.C, off,
                               v.C, C.c_longlong(v.M), st)
                    off += v.C
                if out.needs_stats:
                    self._stats_call(f, out)
            elif k == "gap":
                x, out = op.ins[0], op.out
                self._emit(f, "hg_spatial_mean", self.hdt, L.ptr(x.buf), x.N, x.H, x.W, x.C,
                           C.c_float(1.0 / (x.H * x.W)), None, L.ptr(out.buf), st)
                if out.needs_stats:
                    self._stats_call(f, out)
            elif k == "bcast":
                y, out = op.ins[0], op.out
                self._emit(f, "hg_spatial_broadcast", self.hdt, L.ptr(y.buf), out.N, out.H, out.W, out.C,
                           C.c_float(1.0), None, L.ptr(out.buf), st)
                if out.needs_stats:
                    self._stats_call(f, out)
            elif k == "export":
                v = op.ins[0]
                self._emit(f, "hg_nhwc_to_nchw_f32", self.hdt, L.ptr(v.buf), v.N, v.C, v.H, v.W,
                           L.ptr(self.out_static[self.out_index[id(v)]]), st)
            else:
                raise RuntimeError(f"unknown op {k}")
        # running statistics: one launch for every BatchNorm module of the forward, call sites in order
        self.running_tables = None
        if running:
            mods, sites = [], []
            for bn, lst in running.values():
                mods.append(L.HgBnRunningModule(self._rstat(bn.running_mean).data_ptr(),
                                                self._rstat(bn.running_var).data_ptr(),
                                                bn.num_batches_tracked.data_ptr(), bn.num_features,
                                                L.pad64(bn.num_features), len(sites), len(lst),
                                                float(bn.momentum if bn.momentum is not None else 0.1), 0))
                for stats, cnt in lst:
                    sites.append(L.HgBnRunningSite(stats.data_ptr(), cnt, 0))
            mod_arr = (L.HgBnRunningModule * len(mods))(*mods)
            site_arr = (L.HgBnRunningSite * len(sites))(*sites)
            mod_dev = torch.frombuffer(bytearray(bytes(mod_arr)), dtype=torch.uint8).to(self.device)
            site_dev = torch.frombuffer(bytearray(bytes(site_arr)), dtype=torch.uint8).to(self.device)
            self.running_tables = (mod_dev, site_dev)
            self._cur_lane = 0
            self._emit(f, "hg_bn_update_running", L.ptr(mod_dev), L.ptr(site_dev), len(mods), st).barrier = True

    def _bias_ptr(self, cv, info):
        """Bias vector padded to Cout_p: the live fp32 parameter itself when no padding / cast is needed."""
        if cv.bias is None:
            return None
        if info["mix"] is not None:
            return L.ptr(info["bias"])  # b_eff, refreshed by hg_mix_rows every forward
        if cv.bias.dtype == torch.float32 and L.pad64(cv.out_channels) == cv.out_channels:
            return L.ptr(cv.bias.data)
        info["bias_copy"] = True
        return L.ptr(info["bias"])

    def _stats_call(self, lst, v):
        d = L.HgBnDesc(v.M, v.C, self.hdt, 1e-5, 0, 0)
        self._keep.append(d)
        self._emit(lst, "hg_bn_stats", C.byref(d), L.ptr(v.buf), L.ptr(v.stats), self.stream)

    # ------------------------------------------------------------------------------------------------
    def _lower_backward(self, convs):
        g, st = self.bwd_calls, self.stream
        self._last_writer = {}
        done = set()      # ids of the ops whose backward has been lowered
        deferred = []     # max-pool backwards waiting for the other gradient contributions of their input

        def emit_pool_bwd(op):
            x, G = op.ins[0], op.out.grad
            self._cur_lane = op.lane
            addend, dst = self._grad_target(x)
            self._emit(g, "hg_maxpool2_bwd", self.hdt, L.ptr(x.buf), L.ptr(G), L.ptr(addend), x.N, x.H, x.W,
                       x.C, L.ptr(dst), st)

        for op in reversed(self.b.ops):
            # a deferred max-pool backward goes out as soon as every other consumer of its input has contributed
            # (before the producer of that input is lowered: producers precede all consumers in forward order)
            for dop in list(deferred):
                if all(id(c) in done for c in dop.ins[0].consumers if c is not dop):
                    deferred.remove(dop)
                    emit_pool_bwd(dop)
            done.add(id(op))
            k = op.kind
            self._cur_lane = op.lane
            if k == "export":
                v = op.ins[0]
                if v.requires_grad:
                    addend, out = self._grad_target(v)
                    self._emit(g, "hg_nchw_f32_to_nhwc", self.hdt, L.ptr(self.gout_static[self.out_index[id(v)]]),
                               L.ptr(addend), v.N, v.C, v.H, v.W, L.ptr(out), st)
                continue
            out = op.out
            if k == "conv" and op.attrs["head"] and out.requires_grad:
                # gradient arriving from the loss for this returned heatmap
                addend, dst = self._grad_target(out)
                self._emit(g, "hg_nchw_f32_to_nhwc", self.hdt, L.ptr(self.gout_static[self.out_index[id(out)]]),
                           L.ptr(addend), out.N, out.C, out.H, out.W, L.ptr(dst), st)
            if out.grad is None:
                continue  # nothing downstream of this value reaches a loss (e.g. the last `inter`, quirk Q5)
            G = out.grad
            if k == "conv":
                cv, x, res = op.attrs["conv"], op.ins[0], op.ins[1]
                info = self.conv_info[op.attrs["key"]]
                d = self._conv_desc(cv, x, info["cout"])
                foldop = op.attrs.get("fold")
                if res is not None:
                    self._grad_passthrough(res, G)
                wslot = self._gslot(cv.weight)
                bslot = self._gslot(cv.bias) if op.attrs["use_bias"] else None
                # bias gradient comes for free from the BatchNorm backward when y only feeds a BatchNorm
                bn_only = (len(out.consumers) == 1 and out.consumers[0].kind == "bn" and res is None
                           and not op.attrs["head"] and info["mix"] is None)
                if bn_only:
                    bslot = None
                if info["mix"] is not None and bslot is not None:  # db_eff first, un-mixed at the end
                    bslot = self.packed_arena[info["dbeff_off"]:info["dbeff_off"] + info["cout"]]
                if wslot is not None or bslot is not None:
                    if wslot is None:
                        dwp = None
                    elif info["direct"]:
                        dwp = wslot
                    else:
                        dwp = self.packed_arena[info["goff"]:info["goff"] + info["gsize"]]
                        info["used"] = True
                    # Nothing downstream waits for a weight gradient until the very end of the pass, so the wgrad
                    # kernels leave the critical dgrad -> BN-backward chain: they rotate over dedicated stream lanes
                    # and fill the SMs the latency-bound low-resolution kernels of the main chain leave idle.
                    keep_lane = self._cur_lane
                    if self.wgrad_lanes:
                        self._cur_lane = self.wgrad_lanes[self._wgrad_rr % len(self.wgrad_lanes)]
                        self._wgrad_rr += 1
                    if foldop is not None and foldop.attrs["folded"]:
                        fold, reads = self._bn_fold(foldop)
                        self._emit(g, "hg_conv_wgrad_bn", C.byref(d), C.byref(fold), L.ptr(foldop.ins[0].buf), L.ptr(G),
                                   L.ptr(dwp), L.ptr(bslot), st).tag = self._conv_tag(cv, x) + " +bn"
                    else:
                        self._emit(g, "hg_conv_wgrad", C.byref(d), L.ptr(x.buf), L.ptr(G), L.ptr(dwp), L.ptr(bslot),
                                   st).tag = self._conv_tag(cv, x)
                    self._cur_lane = keep_lane
                    if wslot is not None and not info["direct"]:
                        info["last_wgrad"] = g[-1]
                        info.setdefault("wgrads", []).append(g[-1])
                if x.requires_grad:
                    addend, dst = self._grad_target(x)
                    if foldop is not None:
                        # epilogue = ReLU mask + the two BatchNorm-backward sums; dst holds g = da * [bn(x) > 0]
                        assert addend is None, "a folded BatchNorm output has exactly one consumer"
                        fold, reads = self._bn_fold(foldop)
                        self._emit(g, "hg_conv_dgrad_bn", C.byref(d), C.byref(fold), L.ptr(G), L.ptr(info["wd"]),
                                   L.ptr(foldop.ins[0].buf), L.ptr(dst), L.ptr(foldop.attrs["red"]),
                                   st).tag = self._conv_tag(cv, x) + " +bn"
                    else:
                        self._emit(g, "hg_conv_dgrad", C.byref(d), L.ptr(G), L.ptr(info["wd"]), L.ptr(addend),
                                   L.ptr(dst), st).tag = self._conv_tag(cv, x)
            elif k == "bn":
                bn, x = op.attrs["bn"], op.ins[0]
                d = self._bn_desc(bn, x, op.attrs["relu"])
                red = op.attrs["red"]
                gam, bet = L.ptr(self._p32(bn.weight)), L.ptr(self._p32(bn.bias))
                stats = L.ptr(x.stats) if x.stats is not None else None
                rmean = L.ptr(self._rstat(bn.running_mean))
                rvar = L.ptr(self._rstat(bn.running_var))
                if op.attrs.get("masked"):
                    pass  # the sums were accumulated by the consumer's hg_conv_dgrad_bn; G is already masked
                elif not d.use_running or self._gslot_peek(bn.weight) or self._gslot_peek(bn.bias):
                    self._emit(g, "hg_bn_bwd_reduce", C.byref(d), L.ptr(G), L.ptr(x.buf), stats, gam, bet, rmean, rvar,
                               L.ptr(red), st)
                colsum = None
                prod = x.producer
                if (prod is not None and prod.kind == "conv" and len(x.consumers) == 1 and prod.ins[1] is None
                        and not prod.attrs["head"] and prod.attrs["use_bias"] and prod.attrs["mix"] is None):
                    colsum = self._gslot(prod.attrs["conv"].bias)
                if x.requires_grad:
                    addend, dst = self._grad_target(x)
                else:
                    addend, dst = None, self._scratch(x)
                self._emit(g, "hg_bn_bwd_apply", C.byref(d), L.ptr(G), L.ptr(x.buf), stats, gam, bet,
                           L.ptr(self._rstat(bn.running_mean)),
                           L.ptr(self._rstat(bn.running_var)), L.ptr(red), L.ptr(addend),
                           L.ptr(dst), L.ptr(self._gslot(bn.weight)), L.ptr(self._gslot(bn.bias)), L.ptr(colsum), st)
            elif k == "pool":
                x = op.ins[0]
                if x.requires_grad:
                    # The hourglass input feeds the skip branch (a residual block: its gradient arrives as an identity
                    # pass-through plus a BatchNorm backward) and this pool.  Lowered in reverse op order the pool would
                    # come first and the pass-through would then cost a separate hg_add (201 MB at 64x64, 32 per step):
                    # the max-pool backward has an addend input, so it goes LAST instead and adds into the sum.
                    if _DEFER_POOL_BWD and any(id(c) not in done for c in x.consumers):
                        deferred.append(op)
                    else:
                        emit_pool_bwd(op)
            elif k == "up":
                low, skip = op.ins[0], op.ins[1]
                if skip is not None:
                    self._grad_passthrough(skip, G)
                if low.requires_grad:
                    addend, dst = self._grad_target(low)
                    self._emit(g, "hg_upsample2x_bwd", self.hdt, op.attrs["mode"], L.ptr(G), L.ptr(addend), low.N,
                               low.H, low.W, low.C, L.ptr(dst), st)
            elif k == "add":
                self._grad_passthrough(op.ins[0], G)
                self._grad_passthrough(op.ins[1], G)
            elif k == "cat":
                off = 0
                for v in op.ins:
                    if v.requires_grad:
                        addend, dst = self._grad_target(v)
                        self._emit(g, "hg_channel_copy", self.hdt, L.ptr(G), out.C, off, L.ptr(addend), L.ptr(dst), v.C,
                                   0, v.C, C.c_longlong(v.M), st)
                    off += v.C
            elif k == "gap":
                x = op.ins[0]
                if x.requires_grad:
                    addend, dst = self._grad_target(x)
                    self._emit(g, "hg_spatial_broadcast", self.hdt, L.ptr(G), x.N, x.H, x.W, x.C,
                               C.c_float(1.0 / (x.H * x.W)), L.ptr(addend), L.ptr(dst), st)
            elif k == "bcast":
                y = op.ins[0]
                if y.requires_grad:
                    addend, dst = self._grad_target(y)
                    self._emit(g, "hg_spatial_mean", self.hdt, L.ptr(G), out.N, out.H, out.W, out.C, C.c_float(1.0),
                               L.ptr(addend), L.ptr(dst), st)
            elif k == "stem":
                cv, x = op.attrs["conv"], op.ins[0]
                self._emit(g, "hg_stem_bwd", self.hdt, L.ptr(x.buf), L.ptr(out.buf), L.ptr(G), x.N, x.H, x.W,
                           1 if op.attrs["relu"] else 0, L.ptr(self._gslot(cv.weight)),
                           L.ptr(self._gslot(cv.bias)) if cv.bias is not None else None, st)
        for dop in deferred:   # (inputs without a producer op)
            emit_pool_bwd(dop)
        self._cur_lane = 0
        # graph inputs that want a gradient (sub-module use): NHWC -> NCHW fp32
        self.gin_static = None
        for v, kind in self.b.inputs:
            if kind == "nchw" and v.requires_grad:
                self.gin_static = torch.zeros(v.N, v.C, v.H, v.W, device=self.device, dtype=torch.float32)
                if v.grad is not None:
                    self._emit(g, "hg_nhwc_to_nchw_f32", self.hdt, L.ptr(v.grad), v.N, v.C, v.H, v.W,
                               L.ptr(self.gin_static), st)
        # packed weight gradients -> OIHW slots, right after the LAST call site of each shared weight so that the
        # gradient is final as early as possible (bucketed all-reduce, parallel.py)
        for key in convs:
            info = self.conv_info[key]
            if not info.get("used"):
                continue
            cv = info["conv"]
            k = cv.kernel_size[0]
            d = L.HgConvDesc(1, 1, 1, info["cin_seg"], info["cout"], k, k, 1, 0, 1, self.hdt)
            self._keep.append(d)
            dwp = self.packed_arena[info["goff"]:info["goff"] + info["gsize"]]
            self._pending_writes = []
            slot = self._gslot(cv.weight)
            widx = self.param_index[id(cv.weight)]
            tail = []
            if info["mix"] is None:
                tail.append(_Call("hg_unpack_conv_wgrad_slice", (C.byref(d), L.ptr(dwp), L.ptr(slot), cv.in_channels,
                                                                 info["cin_off"], 1, st)))
            else:  # dW = T^T dW_eff, db = T^T db_eff
                n = info["cout"] * cv.in_channels * k * k
                dweff = self.packed_arena[info["dweff_off"]:info["dweff_off"] + n]
                tail.append(_Call("hg_unpack_conv_wgrad_slice", (C.byref(d), L.ptr(dwp), L.ptr(dweff), cv.in_channels,
                                                                 0, 0, st)))
                tail.append(_Call("hg_mix_rows_rect", (L.ptr(info["mix"]), L.ptr(dweff), L.ptr(slot), info["cout"],
                                                       cv.out_channels, cv.in_channels * k * k, 1, 1, st)))
                if cv.bias is not None and self._gslot(cv.bias) is not None:
                    dbeff = self.packed_arena[info["dbeff_off"]:info["dbeff_off"] + info["cout"]]
                    tail.append(_Call("hg_mix_rows_rect", (L.ptr(info["mix"]), L.ptr(dbeff),
                                                           L.ptr(self._gslot(cv.bias)), info["cout"], cv.out_channels,
                                                           1, 1, 1, st)))
            tail[-1].writes = tuple(self._pending_writes)
            self._pending_writes = []
            # the wgrad calls themselves only touch the packed accumulator, not the OIHW slot
            for call in g:
                if call.name == "hg_conv_wgrad" and widx in call.writes:
                    call.writes = tuple(w for w in call.writes if w != widx)
            # The unpack runs on the lane of the weight's last wgrad call and waits (events) for the call sites that ran
            # on the OTHER wgrad lanes -- not on the main lane behind a barrier over every lane, which made the critical
            # dgrad / BatchNorm chain of the last stack wait ~30 times for all queued low-priority wgrad work to drain.
            pos = g.index(info["last_wgrad"]) + 1
            lane = info["last_wgrad"].lane
            last_per_lane = {}
            for wc in info.get("wgrads", []):
                last_per_lane[wc.lane] = wc
            deps = tuple(wc for ln, wc in last_per_lane.items() if ln != lane)
            for i, c in enumerate(tail):
                c.lane = lane
                c.barrier = _UNPACK_BARRIER and i == 0
                if _UNPACK_BARRIER:
                    c.lane = 0
                elif i == 0:
                    c.deps = deps
                g.insert(pos + i, c)

    def plan_gradient_buckets(self, quantiles=(0.5, 0.97)):
        """Split the backward call list at the points where the given fractions of the gradient volume are final.
        In the weight-shared family every hourglass weight is complete only during the FIRST stack's backward
        ('bucket by last grad accumulation'), the stem's weights at the very end.  Returns the segments
        run_backward executes: (first_call, end_call, element ranges of grad_arena that are final after it)."""
        g = self.bwd_calls
        ready = [-1] * len(self.params)
        for idx, call in enumerate(g):
            for w in call.writes:
                ready[w] = idx
        sizes = [p.numel() for _, p in self.params]
        order = sorted(range(len(sizes)), key=lambda i: ready[i])
        total = float(sum(sizes))
        cuts, acc, qi = [], 0, 0
        for i in order:
            acc += sizes[i]
            while qi < len(quantiles) and acc >= quantiles[qi] * total:
                cuts.append(ready[i] + 1)
                qi += 1
        bounds = sorted(set(c for c in cuts if 0 < c < len(g))) + [len(g)]
        offs = self.grad_offsets
        out, first = [], 0
        for end in bounds:
            ranges = []
            for i in range(len(sizes)):
                if first <= ready[i] < end or (first == 0 and ready[i] < 0):
                    hi = offs[i] + (sizes[i] + 3) // 4 * 4  # slots are padded to 16 B; the padding stays zero
                    if ranges and ranges[-1][1] == offs[i]:
                        ranges[-1][1] = hi
                    else:
                        ranges.append([offs[i], hi])
            out.append((first, end, [tuple(r) for r in ranges]))
            first = end
        self.bwd_segments = out
        return out

    def _scratch(self, v):
        t = self._act(v)
        self._keep.append(t)
        return t

    # ------------------------------------------------------------------------------------------------
    # execution
    # ------------------------------------------------------------------------------------------------
    def _run_calls(self, calls, bwd=False):
        if self.num_lanes > 1 and self.profile_records is None:
            return self._run_calls_lanes(calls, bwd)
        self.stream.value = torch.cuda.current_stream().cuda_stream
        if self.profile_records is not None:
            # per-call CUDA events on the launching stream (bench.py's per-kernel table / roofline)
            for c in calls:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = c.fn(*c.args)
                e1.record()
                if rc != 0:
                    raise RuntimeError(f"libhg_sm100a: {c.name} failed with status {rc}: {L.last_error()}")
                self.profile_records.append((c.name, c.tag, e0, e1))
            return
        for c in calls:
            rc = c.fn(*c.args)
            if rc != 0:
                raise RuntimeError(f"libhg_sm100a: {c.name} failed with status {rc}: {L.last_error()}")

    def _run_calls_lanes(self, calls, bwd=False):
        if os.environ.get("HG_DEBUG_SKIP"):   # timing experiments only: drop some entry points (results are wrong)
            # entries: `entry_point` or `entry_point:tag substring` (e.g. hg_conv_wgrad:@4x4)
            skip = [tuple(e.split(":", 1)) for e in os.environ["HG_DEBUG_SKIP"].split(",")]

            def dropped(c):
                return any(c.name == e[0] and (len(e) == 1 or e[1] in (c.tag or "")) for e in skip)
            calls = [c for c in calls if not dropped(c)]
            for c in calls:
                c.deps = tuple(d for d in c.deps if not dropped(d))
        """Issue the calls on their stream lanes; cross-lane dependencies become event waits (graph edges when
        captured).  Every lane is joined back into the main stream at the end."""
        main = torch.cuda.current_stream()
        if bwd not in self.lane_streams:
            side = [None]
            for lane in range(1, self.num_lanes):
                pr = (_BWD_PRIO[2] if lane in self.wgrad_lanes else _BWD_PRIO[1]) if bwd else 0
                side.append(torch.cuda.Stream(priority=pr))
            self.lane_streams[bwd] = side
        streams = [main] + self.lane_streams[bwd][1:]
        needed = set()
        for c in calls:
            for d in c.deps:
                needed.add(id(d))
        in_list = set(id(c) for c in calls)
        touched = [False] * self.num_lanes
        touched[0] = True
        fork = None
        for c in calls:
            s = streams[c.lane]
            if not touched[c.lane]:
                # First use of a side lane.  Its data dependencies (events recorded on the main lane right after
                # the producing calls) fork it off the main stream at the right point; only a call without any
                # dependency needs an explicit fork from "now".
                if not any(id(d) in in_list and d.event is not None for d in c.deps):
                    fork = torch.cuda.Event()
                    fork.record(main)
                    s.wait_event(fork)
                touched[c.lane] = True
            if c.barrier:
                for l in range(1, self.num_lanes):
                    if touched[l]:
                        e = torch.cuda.Event()
                        e.record(streams[l])
                        s.wait_event(e)
            for d in c.deps:
                if id(d) in in_list and d.event is not None:
                    s.wait_event(d.event)
            self.stream.value = s.cuda_stream
            rc = c.fn(*c.args)
            if rc != 0:
                raise RuntimeError(f"libhg_sm100a: {c.name} failed with status {rc}: {L.last_error()}")
            if id(c) in needed:
                c.event = torch.cuda.Event()
                c.event.record(s)
        for l in range(1, self.num_lanes):
            if touched[l]:
                e = torch.cuda.Event()
                e.record(streams[l])
                main.wait_event(e)

    def _fwd_body(self):
        for p, shadow in self._shadow.values():
            shadow.copy_(p.data)
        for buf, shadow in self._shadow_buf.values():
            shadow.copy_(buf)
        self._refresh_bias()
        self._run_calls(self.fwd_calls)
        if self.running_tables is not None:
            for buf, shadow in self._shadow_buf.values():
                buf.copy_(shadow)

    def _refresh_bias(self):
        # padded fp32 bias vectors (device-to-device copies of the live parameters)
        for cv, info in self._bias_pairs:
            info["bias"][:cv.out_channels].copy_(cv.bias.data)

    def _bwd_body(self, a=0, b=None):
        if a == 0:
            L.zero_(self.grad_arena)      # memset nodes, no fill kernels
            L.zero_(self.packed_arena)
            L.zero_(self.red_arena)
        self._run_calls(self.bwd_calls[a:b], bwd=True)

    def prepare(self):
        self._bias_pairs = [(info["conv"], info) for info in self.conv_info.values()
                            if info["conv"].bias is not None and info.get("bias_copy")]

    def run_forward(self, x):
        inp = self.b.inputs[0][0]
        target = inp.buf if self.b.inputs[0][1] == "image" else self.in_nchw
        target.copy_(x)
        if not hasattr(self, "_bias_pairs"):
            self.prepare()
        if _USE_GRAPHS and self.n_fwd_runs >= 1 and self.profile_records is None:
            if self.fwd_graph is None:
                self.fwd_graph = self._capture(self._fwd_body, _FWD_MAIN_PRIORITY)
            self.fwd_graph.replay()
        else:
            self._fwd_body()
        self.n_fwd_runs += 1
        # the caller owns what it gets (a stock module returns new tensors every call): ONE device-to-device copy of the
        # output arena (a memcpy, no kernel); the returned tensors are views of the copy
        flat = self.out_arena.clone()
        return [flat[o:o + shp[0] * shp[1] * shp[2] * shp[3]].view(shp) for o, shp in self._out_slices]

    def run_backward(self, gouts):
        present = [(gbuf, gouts[i]) for i, gbuf in enumerate(self.gout_static) if i < len(gouts) and gouts[i] is not None]
        if len(present) < len(self.gout_static):
            L.zero_(self.gout_arena)
        if present:   # incoming heatmap gradients: device-to-device memcpys when they are plain fp32 tensors
            if all(g.dtype == torch.float32 and g.is_contiguous() and g.is_cuda for _, g in present):
                for dst, g in present:
                    dst.copy_(g, non_blocking=True)
            else:
                torch._foreach_copy_([p[0] for p in present], [p[1].to(torch.float32) for p in present])
        segments = self.bwd_segments if (self.reducer is not None and self.bwd_segments) else [(0, None, None)]
        for a, b, ranges in segments:
            if _USE_GRAPHS and self.n_fwd_runs >= 2 and self.profile_records is None:
                if (a, b) not in self.bwd_graphs:
                    self.bwd_graphs[(a, b)] = self._capture(lambda a=a, b=b: self._bwd_body(a, b), _BWD_PRIO[0])
                self.bwd_graphs[(a, b)].replay()
            else:
                self._bwd_body(a, b)
            if self.reducer is not None:
                self.reducer.reduce_async(self.grad_arena, ranges)  # overlaps the next segment
        if self.reducer is not None:
            self.reducer.wait()
        grads = []
        flat = self.grad_arena.clone()
        for i, (_, p) in enumerate(self.params):
            n, off = p.numel(), self.grad_offsets[i]
            if self.param_used[i]:
                gp = flat[off:off + n].view(p.shape)
                grads.append(gp if p.dtype == torch.float32 else gp.to(p.dtype))
            else:
                grads.append(None)
        gin = self.gin_static.clone() if self.gin_static is not None else None
        return gin, grads

    def _capture(self, body, priority=0):
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(priority=priority)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            with torch.cuda.graph(graph, stream=s):
                body()
        torch.cuda.current_stream().wait_stream(s)
        return graph

    @property
    def launches_fwd(self):
        return len(self.fwd_calls)

    @property
    def launches_bwd(self):
        return len(self.bwd_calls)


class _PlanFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, *params):
        outs = plan.run_forward(x)
        ctx.plan = plan
        ctx.generation = plan.n_fwd_runs
        ctx.x_needs_grad = x.requires_grad
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        if ctx.generation != ctx.plan.n_fwd_runs:
            raise RuntimeError("backward() of a forward pass whose saved activations were overwritten by a later "
                               "forward of the same module/shape: plans own one static set of buffers (like a CUDA "
                               "graph); run backward before the next forward")
        gin, grads = ctx.plan.run_backward(gouts)
        return (None, gin if ctx.x_needs_grad else None, *grads)


def run_plan(plan, x):
    params = [p for _, p in plan.params]
    if plan.need_bwd and torch.is_grad_enabled():
        outs = _PlanFunction.apply(plan, x, *params)
    else:
        outs = tuple(plan.run_forward(x))
    return list(outs)
