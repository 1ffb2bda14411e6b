"""GPU parity tests of the individual kernels, called through the C ABI (ctypes), against plain PyTorch fp32
ops on the same inputs.  bf16 tolerance: rtol 2e-2 of the tensor's max (north_star); fp32 path: 1e-5."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

DTYPES = [torch.bfloat16, torch.float32]


@pytest.fixture(autouse=True)
def _exact_reference():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def nhwc(t, dtype):
    n, c, h, w = t.shape
    out = torch.zeros(n, h, w, L.pad64(c), device=t.device, dtype=dtype)
    out[..., :c] = t.permute(0, 2, 3, 1).to(dtype)
    return out.contiguous()


def nchw(t, c):
    return t[..., :c].float().permute(0, 3, 1, 2).contiguous()


def tol(dtype):
    return 2e-2 if dtype == torch.bfloat16 else 1e-5


def close(got, ref, rtol, what=""):
    scale = ref.abs().max().item() + 1e-20
    err = (got - ref).abs().max().item()
    assert err <= rtol * scale, f"{what}: max abs err {err:.3e} vs scale {scale:.3e} (rtol {rtol})"


CONV_CASES = [
    # N, H, W, Cin, Cout, k, dil, residual, head
    (2, 16, 16, 64, 64, 1, 1, False, False),
    (2, 16, 16, 256, 128, 1, 1, False, False),
    (2, 64, 64, 128, 128, 3, 1, False, False),
    (2, 64, 64, 128, 256, 1, 1, True, False),
    (4, 4, 4, 128, 128, 3, 1, False, False),
    (3, 8, 8, 128, 128, 3, 1, False, False),
    (1, 128, 128, 64, 64, 3, 1, False, False),
    (1, 128, 128, 64, 128, 1, 1, True, False),
    (2, 32, 32, 256, 256, 1, 1, True, False),
    (2, 64, 64, 256, 16, 1, 1, False, True),
    (2, 64, 64, 16, 256, 1, 1, True, False),
    (2, 64, 64, 256, 17, 1, 1, False, True),
    (2, 64, 64, 256, 38, 1, 1, False, True),
    (2, 4, 4, 256, 256, 3, 6, False, False),
    (5, 32, 32, 128, 128, 3, 1, False, False),   # M not a multiple of the CTA tile at N=5? (5*1024 = 40 tiles) fine
    (1, 8, 8, 256, 128, 1, 1, False, False),     # M = 64 < one tile: TMA clips / zero-fills
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fprop_dgrad_wgrad(case, dtype):
    N, H, W, Cin, Cout, k, dil, res, head = case
    torch.manual_seed(0)
    dev = "cuda"
    pad = dil * (k // 2)
    x = torch.randn(N, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device=dev)
    r = torch.randn(N, Cout, H, W, device=dev) if res else None
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 1, pad, dil, L.hg_dtype(dtype))
    Cin_p, Cout_p = L.pad64(Cin), L.pad64(Cout)
    xq, rq = nhwc(x, dtype), (nhwc(r, dtype) if res else None)
    wf = torch.empty(k * k, Cout_p, Cin_p, device=dev, dtype=dtype)
    wd = torch.empty(k * k, Cin_p, Cout_p, device=dev, dtype=dtype)
    bias_p = torch.zeros(Cout_p, device=dev)
    bias_p[:Cout] = b
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), st)
    y = torch.full((N, H, W, Cout_p), float("nan"), device=dev, dtype=dtype)
    stats = torch.zeros(3 * Cout_p, device=dev)   # {S1, S2, pivot}: pivot 0 = plain sums
    out_nchw = torch.full((N, Cout, H, W), float("nan"), device=dev) if head else None
    L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xq), L.ptr(wf), L.ptr(bias_p), L.ptr(rq), L.ptr(y), L.ptr(stats),
           L.ptr(out_nchw), st)
    xr, wr = nchw(xq, Cin), w.to(dtype).float()
    ref = F.conv2d(xr, wr, b, 1, pad, dil)
    if res:
        ref = ref + nchw(rq, Cout)
    close(nchw(y, Cout), ref, tol(dtype), "fprop")
    if Cout_p > Cout:
        assert y[..., Cout:].float().abs().max().item() == 0.0, "padded output channels must stay zero"
    if head:
        close(out_nchw, ref, 1e-4 if dtype == torch.bfloat16 else 1e-5, "fprop fp32 NCHW copy")
    yv = y[..., :Cout].float()
    close(stats[:Cout], yv.sum((0, 1, 2)), 1e-3, "stats sum")
    close(stats[Cout_p:Cout_p + Cout], (yv * yv).sum((0, 1, 2)), 1e-3, "stats sumsq")

    dy = torch.randn(N, Cout, H, W, device=dev)
    dyq = nhwc(dy, dtype)
    dyr = nchw(dyq, Cout)
    add = torch.randn(N, Cin, H, W, device=dev)
    addq = nhwc(add, dtype)
    dx = torch.full((N, H, W, Cin_p), float("nan"), device=dev, dtype=dtype)
    L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), L.ptr(addq), L.ptr(dx), st)
    ref_dx = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, 1, pad, dil) + nchw(addq, Cin)
    close(nchw(dx, Cin), ref_dx, tol(dtype), "dgrad (+addend)")
    # in-place accumulation (addend == dx), as the plan uses it
    dx2 = addq.clone()
    L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), L.ptr(dx2), L.ptr(dx2), st)
    assert torch.equal(dx2, dx), "in-place dgrad accumulation differs"

    dwp = torch.zeros(k * k, Cout_p, Cin_p, device=dev)
    dbias = torch.zeros(Cout, device=dev)
    for _ in range(2):  # accumulates over call sites
        L.call("hg_conv_wgrad", C.byref(d), L.ptr(xq), L.ptr(dyq), L.ptr(dwp), L.ptr(dbias), st)
    dw = torch.zeros_like(w)
    L.call("hg_unpack_conv_wgrad", C.byref(d), L.ptr(dwp), L.ptr(dw), 0, st)
    ref_dw = 2 * torch.nn.grad.conv2d_weight(xr, w.shape, dyr, 1, pad, dil)
    close(dw, ref_dw, 1e-3 if dtype == torch.bfloat16 else 2e-5, "wgrad")
    close(dbias, 2 * dyr.sum((0, 2, 3)), 1e-3, "dbias")


P1_CASES = [c for c in CONV_CASES if c[5] == 1 and not c[8] and (c[0] * c[1] * c[2]) % 128 == 0] + [
    (12, 64, 64, 256, 128, 1, 1, False, False),   # 384 units: the size class the dispatcher picks by itself
    (10, 64, 64, 128, 256, 1, 1, True, False),    # Np = 256: two 128-column parts per 128-pixel tile, weights resident
    (3, 32, 32, 256, 256, 1, 1, True, False),     # 24 units < 148 CTAs
    (19, 32, 32, 256, 128, 1, 1, False, False),   # 152 units on 76 CTAs ... 148: ranges of 1 / 2 units
    (5, 16, 16, 128, 64, 1, 1, True, False),      # 10 units, Np = 64 (one 64-column panel)
    (37, 16, 16, 64, 256, 1, 1, False, False),    # one 64-channel slice, 74 units
]


def _persist(min_units=512, p1=1, p3=1):
    L.call("hg_set_option", b"persist_1x1", p1)
    L.call("hg_set_option", b"persist_3x3", p3)
    L.call("hg_set_option", b"persist_min_units", min_units)


@pytest.mark.parametrize("case", P1_CASES)
def test_persistent_pointwise_kernel(case):
    """conv_persist_kernel on 1x1 convolutions (one resident CTA per SM, weights resident in shared memory, two TMEM
    accumulator sets) forced for every size (persist_min_units = 1): same checks as the tile-per-CTA kernel -- fprop
    (+bias, +residual, statistics), dgrad (+addend, in place), against PyTorch fp32 on bf16-rounded operands."""
    _persist(1)
    try:
        n0 = L.load().hg_launch_count()
        test_conv_fprop_dgrad_wgrad(case, torch.bfloat16)
        assert L.load().hg_launch_count() > n0
    finally:
        _persist(512)


P3_CASES = [
    (2, 64, 64, 128, 128, 3, 1, False, False),    # 64 units on 32 CTAs
    (19, 64, 64, 128, 128, 3, 1, False, False),   # 608 units on 148 CTAs: ranges of 4 / 5 units (odd: 128-pixel tiles)
    (5, 32, 32, 128, 128, 3, 1, True, False),     # + residual
    (40, 32, 32, 128, 128, 3, 1, False, False),   # 320 units: ranges of 2 / 3 units, an image is 8 units
    (3, 128, 128, 64, 64, 3, 1, False, False),    # one 64-channel slice, 64-channel N tile, 2-row tiles
    (9, 16, 16, 128, 64, 3, 1, False, False),     # a tile is a whole 16x16 image
    (3, 32, 32, 256, 128, 3, 1, False, False),    # four 64-channel slices
]


@pytest.mark.parametrize("case", P3_CASES)
def test_persistent_3x3_kernel(case):
    """conv3x3_persist_kernel (256-pixel tiles, one activation box per three taps, two TMEM accumulator sets, one CTA
    per SM over a contiguous unit range) forced for every size: same checks as the tile-per-CTA kernel -- fprop (+bias,
    +residual, statistics), dgrad (+addend, in place) against PyTorch fp32 on bf16-rounded operands."""
    _persist(1, p3=2)
    try:
        n0 = L.load().hg_launch_count()
        test_conv_fprop_dgrad_wgrad(case, torch.bfloat16)
        assert L.load().hg_launch_count() > n0
    finally:
        _persist(512)


@pytest.mark.parametrize("case", [(19, 64, 64, 128, 128, 3, 1, False, False), (12, 64, 64, 256, 128, 1, 1, False, False),
                                  (10, 64, 64, 128, 256, 1, 1, True, False), (40, 32, 32, 128, 128, 3, 1, False, False)])
def test_persistent_kernel_dynamic_tile_order(case):
    """persist_dynamic = 1: tiles handed out by a grid-wide atomic counter through the shared-memory tile queue (off by
    default: measured slower) -- same checks as the static unit ranges, twice (the counter re-arms itself)."""
    _persist(1, p3=2)
    L.call("hg_set_option", b"persist_dynamic", 1)
    try:
        for _ in range(2):
            test_conv_fprop_dgrad_wgrad(case, torch.bfloat16)
    finally:
        L.call("hg_set_option", b"persist_dynamic", 0)
        _persist(512)


def test_persistent_3x3_kernel_matches_tile_kernel():
    """Same 3x3 convolution through the persistent kernel and through conv_gemm_kernel (persist_3x3 = 0): the outputs
    differ only by the fp32 accumulation order of the 18 K blocks (a bf16 rounding step at most), the statistics by 1e-4."""
    torch.manual_seed(11)
    dev, dtype = "cuda", torch.bfloat16
    N, H, W, Cc = 32, 32, 32, 128
    d = L.HgConvDesc(N, H, W, Cc, Cc, 3, 3, 1, 1, 1, L.HG_BF16)
    x = torch.randn(N, H, W, Cc, device=dev).to(dtype)
    w = torch.randn(Cc, Cc, 3, 3, device=dev) / (Cc * 9) ** 0.5
    wf = torch.empty(9, Cc, Cc, device=dev, dtype=dtype)
    wd = torch.empty(9, Cc, Cc, device=dev, dtype=dtype)
    bias = torch.randn(Cc, device=dev)
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), st)
    outs = []
    for on in (1, 0):
        _persist(1, p3=on)
        try:
            n0 = L.load().hg_launch_count()
            y = torch.empty(N, H, W, Cc, device=dev, dtype=dtype)
            stats = torch.zeros(3 * Cc, device=dev)
            L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(x), L.ptr(wf), L.ptr(bias), None, L.ptr(y), L.ptr(stats), None, st)
            outs.append((y.float(), stats.clone()))
        finally:
            _persist(512)
    (y1, s1), (y0, s0) = outs
    assert (y1 - y0).abs().max().item() <= 2 ** -7 * y0.abs().max().item()
    assert (y1 != y0).float().mean().item() < 0.02
    close(s1[:2 * Cc], s0[:2 * Cc], 1e-4, "statistics")


@pytest.mark.parametrize("case", [(2, 64, 64, 128, 128, 3, 1, False, True, False), (19, 64, 64, 128, 128, 3, 1, False, True, False),
                                  (40, 32, 32, 128, 128, 3, 1, False, True, False), (2, 32, 32, 64, 64, 3, 1, False, False, False),
                                  (5, 16, 16, 128, 128, 3, 1, False, True, True)])
def test_persistent_3x3_kernel_masked_dgrad(case):
    """The ReLU-mask / BatchNorm-backward-sums epilogue (hg_conv_dgrad_bn) of the persistent 3x3 kernel: bit-identical
    masked gradient, same sums as hg_bn_bwd_reduce."""
    N, H, W, Cin, Cout, k, dil, res, relu, eval_mode = case
    torch.manual_seed(4)
    dev, dtype = "cuda", torch.bfloat16
    M = N * H * W
    xq = nhwc(torch.randn(N, Cin, H, W, device=dev) * 1.7 + 0.3, dtype)
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    gamma = torch.rand(Cin, device=dev) + 0.5
    beta = torch.randn(Cin, device=dev) * 0.3
    rmean = torch.randn(Cin, device=dev) * 0.2 + 0.3
    rvar = torch.rand(Cin, device=dev) * 2 + 1.5
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 1, 1, 1, L.HG_BF16)
    Cin_p, Cout_p = L.pad64(Cin), L.pad64(Cout)
    wf = torch.empty(k * k, Cout_p, Cin_p, device=dev, dtype=dtype)
    wd = torch.empty(k * k, Cin_p, Cout_p, device=dev, dtype=dtype)
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), st)
    bnd = L.HgBnDesc(M, Cin, L.HG_BF16, 1e-5, 1 if relu else 0, 1 if eval_mode else 0)
    xstats = torch.zeros(3 * Cin_p, device=dev)
    L.call("hg_bn_stats", C.byref(bnd), L.ptr(xq), L.ptr(xstats), st)
    fold = L.HgBnFold(xstats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rmean.data_ptr(), rvar.data_ptr(), 1e-5,
                      1 if relu else 0, 1 if eval_mode else 0, 0)
    a = torch.empty_like(xq)
    L.call("hg_bn_apply", C.byref(bnd), L.ptr(xq), L.ptr(xstats), L.ptr(gamma), L.ptr(beta), L.ptr(rmean),
           L.ptr(rvar), L.ptr(a), st)
    dyq = nhwc(torch.randn(N, Cout, H, W, device=dev), dtype)
    _persist(1, p3=2)
    try:
        da = torch.empty(N, H, W, Cin_p, device=dev, dtype=dtype)
        L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), None, L.ptr(da), st)
        g1 = torch.full((N, H, W, Cin_p), float("nan"), device=dev, dtype=dtype)
        red1 = torch.zeros(2 * Cin_p, device=dev)
        L.call("hg_conv_dgrad_bn", C.byref(d), C.byref(fold), L.ptr(dyq), L.ptr(wd), L.ptr(xq), L.ptr(g1), L.ptr(red1), st)
    finally:
        _persist(512)
    ref_da = torch.nn.grad.conv2d_input((N, Cin, H, W), w.to(dtype).float(), nchw(dyq, Cout), 1, 1, 1)
    close(nchw(da, Cin), ref_da, 2e-2, "dgrad vs torch")
    red0 = torch.zeros(2 * Cin_p, device=dev)
    L.call("hg_bn_bwd_reduce", C.byref(bnd), L.ptr(da), L.ptr(xq), L.ptr(xstats), L.ptr(gamma), L.ptr(beta),
           L.ptr(rmean), L.ptr(rvar), L.ptr(red0), st)
    mask = (a.float() > 0) if relu else torch.ones_like(a, dtype=torch.bool)
    g0 = torch.where(mask, da.float(), torch.zeros((), device=dev))
    assert torch.equal(g1.float(), g0), f"masked dgrad differs: {(g1.float() - g0).abs().max()}"
    scale = red0.abs().max().item()
    assert (red1 - red0).abs().max().item() <= 2e-3 * scale, ((red1 - red0).abs().max().item(), scale)


def test_batchnorm_shifted_statistics_survive_large_means():
    """Statistics slots hold sums of (x - pivot) (include/hg_sm100a.h): with the pivot near the channel mean -- the plan
    uses the consuming BatchNorm's running mean -- the variance of a channel whose mean is 1000x its spread is exact to
    fp32 rounding, where the plain one-pass E[x^2] - E[x]^2 (pivot 0) has lost it entirely.  fp32 path, 32768 rows."""
    torch.manual_seed(5)
    dev, Cc, M = "cuda", 64, 32768
    mean = torch.linspace(50.0, 400.0, Cc, device=dev)
    std = torch.linspace(0.05, 0.4, Cc, device=dev)
    x = (torch.randn(M, Cc, device=dev) * std + mean).contiguous()
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    d = L.HgBnDesc(M, Cc, L.HG_F32, 1e-5, 0, 0)
    st = L.stream_ptr()
    ref = F.batch_norm(x.double().t()[None], None, None, None, None, True, 0.1, 1e-5)[0].t()
    errs = {}
    for name, pivot in (("plain", torch.zeros(Cc, device=dev)), ("shifted", mean + 0.3 * std)):
        slot = torch.zeros(3 * Cc, device=dev)
        slots = (L.HgBnStatsSlot * 1)(L.HgBnStatsSlot(slot.data_ptr(), pivot.data_ptr(), Cc, Cc))
        sdev = torch.frombuffer(bytearray(bytes(slots)), dtype=torch.uint8).cuda()
        slot.fill_(float("nan"))
        L.call("hg_bn_prepare_stats", L.ptr(sdev), 1, st)
        assert torch.equal(slot[2 * Cc:], pivot) and float(slot[:2 * Cc].abs().max()) == 0.0
        L.call("hg_bn_stats", C.byref(d), L.ptr(x), L.ptr(slot), st)
        y = torch.empty_like(x)
        L.call("hg_bn_apply", C.byref(d), L.ptr(x), L.ptr(slot), L.ptr(gamma), L.ptr(beta), None, None, L.ptr(y), st)
        errs[name] = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    # the input itself only carries ~1e-7 * mean / std ~ 1e-4 of relative precision per element in fp32
    assert errs["shifted"] <= 2e-3, errs
    assert errs["plain"] > 10 * errs["shifted"], errs   # what the old formula did to these channels


@pytest.mark.parametrize("case", [(12, 64, 64, 256, 128, 1, 1, False, False), (10, 64, 64, 128, 256, 1, 1, True, False),
                                  (4, 64, 64, 128, 128, 3, 1, False, False), (16, 32, 32, 128, 128, 3, 1, False, False),
                                  (3, 128, 128, 64, 64, 3, 1, False, False)])
def test_wgrad_128_pixel_k_blocks(case):
    """conv_wgrad_kernel with 128-pixel K blocks (the default on the large maps: half as many TMA operations per byte)
    and with 64-pixel K blocks (hg_set_option("wgrad_kpx", 64)): same weight / bias gradients."""
    for kpx in (128, 64):
        L.call("hg_set_option", b"wgrad_kpx", kpx)
        try:
            test_conv_fprop_dgrad_wgrad(case, torch.bfloat16)
        finally:
            L.call("hg_set_option", b"wgrad_kpx", 128)


STRIDE2_CASES = [
    # N, H, W (input), Cin, Cout, k
    (2, 64, 64, 128, 128, 3),     # try_with_aspp_remove_max_pool.py:176 (conv2 of a stride-2 block)
    (2, 64, 64, 256, 256, 1),     # downsaple[0] of the same block (:186), bias-free in the reference
    (3, 128, 128, 64, 64, 3),     # residual1 = ResidualBlock(64, 128, stride=2) (:263)
    (2, 128, 128, 64, 128, 1),
    (4, 8, 8, 128, 128, 3),       # deepest level: 8x8 -> 4x4
    (2, 8, 8, 256, 256, 1),
    (1, 256, 256, 64, 64, 3),     # widest input the stride-2 box takes (output 128 wide)
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", STRIDE2_CASES)
def test_conv_stride2_fprop_dgrad_wgrad(case, dtype):
    """Stride-2 convolutions of the Q4 blocks (try_with_aspp_remove_max_pool.py:165-201, train.py:411-447): on the bf16
    path they run on the tensor-core kernels (element-strided TMA box for fprop / wgrad, one launch per input parity
    class for dgrad); hg_conv_tc_eligible says so and allow_ref_conv stays off."""
    N, H, W, Cin, Cout, k = case
    torch.manual_seed(1)
    dev = "cuda"
    pad = k // 2
    Ho, Wo = H // 2, W // 2
    x = torch.randn(N, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device=dev)
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 2, pad, 1, L.hg_dtype(dtype))
    if dtype == torch.bfloat16:
        assert L.load().hg_conv_tc_eligible(C.byref(d)) == 1
        assert L.load().hg_conv_fold_eligible(C.byref(d)) == 0
    Cin_p, Cout_p = L.pad64(Cin), L.pad64(Cout)
    xq = nhwc(x, dtype)
    wf = torch.empty(k * k, Cout_p, Cin_p, device=dev, dtype=dtype)
    wd = torch.empty(k * k, Cin_p, Cout_p, device=dev, dtype=dtype)
    bias_p = torch.zeros(Cout_p, device=dev)
    bias_p[:Cout] = b
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), st)
    y = torch.full((N, Ho, Wo, Cout_p), float("nan"), device=dev, dtype=dtype)
    stats = torch.zeros(3 * Cout_p, device=dev)
    L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xq), L.ptr(wf), L.ptr(bias_p), None, L.ptr(y), L.ptr(stats), None, st)
    xr, wr = nchw(xq, Cin), w.to(dtype).float()
    ref = F.conv2d(xr, wr, b, 2, pad, 1)
    close(nchw(y, Cout), ref, tol(dtype), "fprop stride 2")
    yv = y[..., :Cout].float()
    close(stats[:Cout], yv.sum((0, 1, 2)), 1e-3, "stats sum")
    close(stats[Cout_p:Cout_p + Cout], (yv * yv).sum((0, 1, 2)), 1e-3, "stats sumsq")

    dy = torch.randn(N, Cout, Ho, Wo, device=dev)
    dyq = nhwc(dy, dtype)
    dyr = nchw(dyq, Cout)
    ref_dx0 = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, 2, pad, 1)
    dx = torch.full((N, H, W, Cin_p), float("nan"), device=dev, dtype=dtype)
    L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), None, L.ptr(dx), st)
    close(nchw(dx, Cin), ref_dx0, tol(dtype), "dgrad stride 2")
    add = torch.randn(N, Cin, H, W, device=dev)
    addq = nhwc(add, dtype)
    dx.fill_(float("nan"))
    L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), L.ptr(addq), L.ptr(dx), st)
    close(nchw(dx, Cin), ref_dx0 + nchw(addq, Cin), tol(dtype), "dgrad stride 2 (+addend)")
    dx2 = addq.clone()
    L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), L.ptr(dx2), L.ptr(dx2), st)
    assert torch.equal(dx2, dx), "in-place strided dgrad accumulation differs"

    dwp = torch.zeros(k * k, Cout_p, Cin_p, device=dev)
    dbias = torch.zeros(Cout, device=dev)
    for _ in range(2):
        L.call("hg_conv_wgrad", C.byref(d), L.ptr(xq), L.ptr(dyq), L.ptr(dwp), L.ptr(dbias), st)
    dw = torch.zeros_like(w)
    L.call("hg_unpack_conv_wgrad", C.byref(d), L.ptr(dwp), L.ptr(dw), 0, st)
    ref_dw = 2 * torch.nn.grad.conv2d_weight(xr, w.shape, dyr, 2, pad, 1)
    close(dw, ref_dw, 1e-3 if dtype == torch.bfloat16 else 2e-5, "wgrad stride 2")
    close(dbias, 2 * dyr.sum((0, 2, 3)), 1e-3, "dbias")


def test_bf16_conv_outside_tensor_core_geometry_fails_loudly():
    """No silent CUDA-core route on the bf16 path: an odd map is HG_ERR_UNSUPPORTED unless allow_ref_conv is set."""
    dev, dtype = "cuda", torch.bfloat16
    N, H, W, Cin, Cout, k = 1, 24, 24, 64, 64, 3
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 1, 1, 1, L.hg_dtype(dtype))
    assert L.load().hg_conv_tc_eligible(C.byref(d)) == 0
    x = torch.randn(N, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    xq = nhwc(x, dtype)
    wf = torch.empty(k * k, 64, 64, device=dev, dtype=dtype)
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), None, st)
    y = torch.zeros(N, H, W, 64, device=dev, dtype=dtype)
    with pytest.raises(RuntimeError, match="allow_ref_conv"):
        L.call("hg_conv_fprop", C.byref(d), L.ptr(xq), L.ptr(wf), None, None, L.ptr(y), None, st)
    L.call("hg_set_option", b"allow_ref_conv", 1)
    try:
        L.call("hg_conv_fprop", C.byref(d), L.ptr(xq), L.ptr(wf), None, None, L.ptr(y), None, st)
        close(nchw(y, Cout), F.conv2d(nchw(xq, Cin), w.to(dtype).float(), None, 1, 1, 1), 2e-2, "ref conv")
    finally:
        L.call("hg_set_option", b"allow_ref_conv", 0)


FOLD_CASES = [
    # N, H, W, Cin, Cout, k, dil, residual, relu, eval_mode
    (2, 64, 64, 128, 128, 3, 1, False, True, False),
    (2, 16, 16, 256, 128, 1, 1, False, True, False),
    (4, 4, 4, 128, 128, 3, 1, False, True, False),
    (3, 8, 8, 128, 256, 1, 1, True, True, False),
    (2, 32, 32, 64, 64, 3, 1, False, False, False),
    (2, 32, 32, 256, 256, 1, 1, True, True, True),
    (5, 16, 16, 128, 128, 3, 1, False, True, True),
    (2, 4, 4, 256, 256, 3, 6, False, True, False),   # dilated taps that only touch padding
    (1, 8, 8, 256, 16, 1, 1, False, True, False),    # M = 64 < one tile, Cout padded to 64
    (2, 128, 128, 64, 128, 1, 1, False, True, False),
]


@pytest.mark.parametrize("case", FOLD_CASES)
def test_conv_with_folded_batchnorm(case):
    """hg_conv_fprop_bn / hg_conv_wgrad_bn / hg_conv_dgrad_bn against the un-fused kernels of the same library
    (hg_bn_apply -> conv, dgrad -> hg_bn_bwd_reduce) and against PyTorch fp32 ops."""
    N, H, W, Cin, Cout, k, dil, res, relu, eval_mode = case
    torch.manual_seed(3)
    dev, dtype = "cuda", torch.bfloat16
    pad = dil * (k // 2)
    M = N * H * W
    x = torch.randn(N, Cin, H, W, device=dev) * 1.7 + 0.3
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device=dev)
    gamma = torch.rand(Cin, device=dev) + 0.5
    beta = torch.randn(Cin, device=dev) * 0.3
    rmean = torch.randn(Cin, device=dev) * 0.2 + 0.3
    rvar = torch.rand(Cin, device=dev) * 2 + 1.5
    r = torch.randn(N, Cout, H, W, device=dev) if res else None
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 1, pad, dil, L.HG_BF16)
    assert L.load().hg_conv_tc_eligible(C.byref(d)) == 1
    Cin_p, Cout_p = L.pad64(Cin), L.pad64(Cout)
    xq, rq = nhwc(x, dtype), (nhwc(r, dtype) if res else None)
    wf = torch.empty(k * k, Cout_p, Cin_p, device=dev, dtype=dtype)
    wd = torch.empty(k * k, Cin_p, Cout_p, device=dev, dtype=dtype)
    bias_p = torch.zeros(Cout_p, device=dev)
    bias_p[:Cout] = b
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), st)
    bnd = L.HgBnDesc(M, Cin, L.HG_BF16, 1e-5, 1 if relu else 0, 1 if eval_mode else 0)
    xstats = torch.zeros(3 * Cin_p, device=dev)
    L.call("hg_bn_stats", C.byref(bnd), L.ptr(xq), L.ptr(xstats), st)
    fold = L.HgBnFold(xstats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rmean.data_ptr(), rvar.data_ptr(), 1e-5,
                      1 if relu else 0, 1 if eval_mode else 0, 0)
    # ---- un-fused path of the same library
    a = torch.empty_like(xq)
    L.call("hg_bn_apply", C.byref(bnd), L.ptr(xq), L.ptr(xstats), L.ptr(gamma), L.ptr(beta), L.ptr(rmean),
           L.ptr(rvar), L.ptr(a), st)
    y0 = torch.empty(N, H, W, Cout_p, device=dev, dtype=dtype)
    s0 = torch.zeros(3 * Cout_p, device=dev)
    L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(a), L.ptr(wf), L.ptr(bias_p), L.ptr(rq), L.ptr(y0), L.ptr(s0), None, st)
    # ---- fused forward
    y1 = torch.full((N, H, W, Cout_p), float("nan"), device=dev, dtype=dtype)
    s1 = torch.zeros(3 * Cout_p, device=dev)
    L.call("hg_conv_fprop_bn", C.byref(d), C.byref(fold), L.ptr(xq), L.ptr(wf), L.ptr(bias_p), L.ptr(rq), L.ptr(y1),
           L.ptr(s1), None, st)
    assert torch.equal(y1, y0), f"fused fprop differs from bn_apply -> conv: {(y1.float() - y0.float()).abs().max()}"
    close(s1, s0, 1e-5, "stats of the fused fprop")
    # ---- PyTorch fp32 reference on the same bf16-representable inputs
    xr = nchw(xq, Cin)
    if eval_mode:
        ar = F.batch_norm(xr, rmean, rvar, gamma, beta, False, 0.0, 1e-5)
    else:
        ar = F.batch_norm(xr, None, None, gamma, beta, True, 0.0, 1e-5)
    if relu:
        ar = F.relu(ar)
    ref = F.conv2d(ar, w.to(dtype).float(), b, 1, pad, dil)
    if res:
        ref = ref + nchw(rq, Cout)
    close(nchw(y1, Cout), ref, 2e-2, "fused fprop vs torch")
    # ---- wgrad
    dy = torch.randn(N, Cout, H, W, device=dev)
    dyq = nhwc(dy, dtype)
    dw0 = torch.zeros(k * k, Cout_p, Cin_p, device=dev)
    dw1 = torch.zeros(k * k, Cout_p, Cin_p, device=dev)
    db0, db1 = torch.zeros(Cout, device=dev), torch.zeros(Cout, device=dev)
    L.call("hg_conv_wgrad", C.byref(d), L.ptr(a), L.ptr(dyq), L.ptr(dw0), L.ptr(db0), st)
    L.call("hg_conv_wgrad_bn", C.byref(d), C.byref(fold), L.ptr(xq), L.ptr(dyq), L.ptr(dw1), L.ptr(db1), st)
    close(dw1, dw0, 1e-5, "fused wgrad vs wgrad on the materialised activation")
    close(db1, db0, 1e-5, "dbias")
    # ---- dgrad with ReLU mask + BatchNorm-backward sums
    da = torch.empty(N, H, W, Cin_p, device=dev, dtype=dtype)
    L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), None, L.ptr(da), st)
    red0 = torch.zeros(2 * Cin_p, device=dev)
    L.call("hg_bn_bwd_reduce", C.byref(bnd), L.ptr(da), L.ptr(xq), L.ptr(xstats), L.ptr(gamma), L.ptr(beta),
           L.ptr(rmean), L.ptr(rvar), L.ptr(red0), st)
    g1 = torch.full((N, H, W, Cin_p), float("nan"), device=dev, dtype=dtype)
    red1 = torch.zeros(2 * Cin_p, device=dev)
    L.call("hg_conv_dgrad_bn", C.byref(d), C.byref(fold), L.ptr(dyq), L.ptr(wd), L.ptr(xq), L.ptr(g1), L.ptr(red1), st)
    mask = (a.float() > 0) if relu else torch.ones_like(a, dtype=torch.bool)
    g0 = torch.where(mask, da.float(), torch.zeros((), device=dev))
    assert torch.equal(g1.float(), g0), f"masked dgrad differs: {(g1.float() - g0).abs().max()}"
    scale = red0.abs().max().item()
    assert (red1 - red0).abs().max().item() <= 2e-3 * scale, ((red1 - red0).abs().max().item(), scale)
    # finishing with hg_bn_bwd_apply(da = g) gives the same dx as the un-fused pair
    dx0 = torch.empty_like(xq)
    dx1 = torch.empty_like(xq)
    for dsrc, redsrc, dst in ((da, red0, dx0), (g1, red1, dx1)):
        L.call("hg_bn_bwd_apply", C.byref(bnd), L.ptr(dsrc), L.ptr(xq), L.ptr(xstats), L.ptr(gamma), L.ptr(beta),
               L.ptr(rmean), L.ptr(rvar), L.ptr(redsrc), None, L.ptr(dx1 if dst is dx1 else dx0), None, None, None, st)
    close(dx1.float(), dx0.float(), 1e-2, "bn backward from the fused sums")


@pytest.mark.parametrize("case", [c for c in FOLD_CASES if c[5] == 1 and (c[0] * c[1] * c[2]) % 128 == 0] +
                         [(10, 64, 64, 256, 128, 1, 1, False, True, False), (10, 64, 64, 128, 256, 1, 1, False, True, False),
                          (19, 32, 32, 256, 256, 1, 1, False, True, False)])
def test_persistent_pointwise_kernel_masked_dgrad(case):
    """The ReLU-mask / BatchNorm-backward-sums epilogue (hg_conv_dgrad_bn) of the persistent kernel on 1x1 convolutions,
    forced for every size: bit-identical masked gradient, same sums as hg_bn_bwd_reduce."""
    _persist(1)
    try:
        test_conv_with_folded_batchnorm(case)
    finally:
        _persist(512)


def test_folded_entry_points_reject_unsupported_geometry():
    d = L.HgConvDesc(2, 32, 32, 128, 128, 3, 3, 2, 1, 1, L.HG_BF16)  # stride 2: tensor cores yes, BatchNorm folding no
    assert L.load().hg_conv_tc_eligible(C.byref(d)) == 1 and L.load().hg_conv_fold_eligible(C.byref(d)) == 0
    t = torch.zeros(4096, device="cuda")
    fold = L.HgBnFold(t.data_ptr(), t.data_ptr(), t.data_ptr(), None, None, 1e-5, 1, 0, 0)
    rc = L.load().hg_conv_fprop_bn(C.byref(d), C.byref(fold), L.ptr(t), L.ptr(t), None, None, L.ptr(t), None, None,
                                   L.stream_ptr())
    assert rc == -2 and "tensor-core" in L.last_error()


def test_conv_tensor_core_matches_cuda_core_kernel():
    """The tcgen05 kernel and the FFMA kernel compute the same bf16 convolution."""
    torch.manual_seed(1)
    N, H, W, Cin, Cout, k = 2, 32, 32, 128, 128, 3
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 1, 1, 1, L.HG_BF16)
    x = nhwc(torch.randn(N, Cin, H, W, device="cuda"), torch.bfloat16)
    w = torch.randn(Cout, Cin, k, k, device="cuda") / 34.0
    wf = torch.empty(9, Cout, Cin, device="cuda", dtype=torch.bfloat16)
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), None, st)
    outs = []
    try:
        for force in (0, 1):
            L.call("hg_set_option", b"force_ref_conv", force)
            y = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.bfloat16)
            L.call("hg_conv_fprop", C.byref(d), L.ptr(x), L.ptr(wf), None, None, L.ptr(y), None, st)
            outs.append(y.float())
    finally:
        L.call("hg_set_option", b"force_ref_conv", 0)
    close(outs[0], outs[1], 1e-2, "tc vs ref")


def test_conv_rejects_bad_descriptor():
    d = L.HgConvDesc(0, 8, 8, 64, 64, 1, 1, 1, 0, 1, L.HG_BF16)
    t = torch.zeros(64, device="cuda")
    rc = L.load().hg_conv_fprop(C.byref(d), L.ptr(t), L.ptr(t), None, None, L.ptr(t), None, L.stream_ptr())
    assert rc == -1 and "non-positive" in L.last_error()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 64, 32, 32), (3, 128, 8, 8), (2, 256, 4, 4), (2, 16, 16, 16)])
@pytest.mark.parametrize("relu", [True, False])
def test_batchnorm_train_fwd_bwd(shape, dtype, relu):
    N, Cc, H, W = shape
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(N, Cc, H, W, device=dev) * 1.7 + 0.4
    gamma = torch.rand(Cc, device=dev) + 0.5
    beta = torch.randn(Cc, device=dev) * 0.3
    xq = nhwc(x, dtype)
    Cp = L.pad64(Cc)
    M = N * H * W
    d = L.HgBnDesc(M, Cc, L.hg_dtype(dtype), 1e-5, 1 if relu else 0, 0)
    st = L.stream_ptr()
    stats = torch.zeros(3 * Cp, device=dev)
    L.call("hg_bn_stats", C.byref(d), L.ptr(xq), L.ptr(stats), st)
    y = torch.empty_like(xq)
    L.call("hg_bn_apply", C.byref(d), L.ptr(xq), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None, L.ptr(y), st)
    xr = nchw(xq, Cc).requires_grad_(True)
    g_ref, b_ref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.batch_norm(xr, None, None, g_ref, b_ref, True, 0.1, 1e-5)
    close(nchw(y, Cc), (F.relu(ref) if relu else ref).detach(), tol(dtype), "bn fwd")
    if relu:  # the ReLU mask of the reference is the kernel's own (an element at 0 +- 1 ulp may go either way)
        ref = ref * (nchw(y, Cc) > 0).float()
    da = torch.randn(N, Cc, H, W, device=dev)
    daq = nhwc(da, dtype)
    ref.backward(nchw(daq, Cc))
    red = torch.zeros(2 * Cp, device=dev)
    L.call("hg_bn_bwd_reduce", C.byref(d), L.ptr(daq), L.ptr(xq), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None, L.ptr(red), st)
    addq = nhwc(torch.randn(N, Cc, H, W, device=dev), dtype)
    dx = torch.empty_like(xq)
    dgamma, dbeta, colsum = (torch.zeros(Cc, device=dev) for _ in range(3))
    L.call("hg_bn_bwd_apply", C.byref(d), L.ptr(daq), L.ptr(xq), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None,
           L.ptr(red), L.ptr(addq), L.ptr(dx), L.ptr(dgamma), L.ptr(dbeta), L.ptr(colsum), st)
    t = 3e-2 if dtype == torch.bfloat16 else 2e-4
    close(nchw(dx, Cc), xr.grad + nchw(addq, Cc), t, "bn dx")
    close(dgamma, g_ref.grad, 1e-3 if dtype == torch.float32 else 2e-2, "bn dgamma")
    close(dbeta, b_ref.grad, 1e-3 if dtype == torch.float32 else 2e-2, "bn dbeta")
    assert colsum.abs().max().item() <= 1e-2 * (xr.grad.abs().sum((0, 2, 3)).max().item() + 1e-6), "sum(dx) ~ 0"


@pytest.mark.parametrize("dtype", DTYPES)
def test_batchnorm_eval_backward(dtype):
    """eval(): running statistics are constants -> dx = gamma*invstd*g, dgamma = sum g*xhat(running), dbeta = sum g."""
    torch.manual_seed(1)
    dev = "cuda"
    N, Cc, H, W = 3, 128, 8, 8
    bn = torch.nn.BatchNorm2d(Cc).to(dev).eval()
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2.0)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_(0, 0.3)
    x = torch.randn(N, Cc, H, W, device=dev)
    da = torch.randn(N, Cc, H, W, device=dev)
    xq, daq = nhwc(x, dtype), nhwc(da, dtype)
    xr = nchw(xq, Cc).requires_grad_(True)
    F.relu(bn(xr)).backward(nchw(daq, Cc))
    st = L.stream_ptr()
    d = L.HgBnDesc(N * H * W, Cc, L.hg_dtype(dtype), 1e-5, 1, 1)
    red = torch.zeros(2 * Cc, device=dev)
    dx = torch.empty_like(xq)
    dg, db = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    L.call("hg_bn_bwd_reduce", C.byref(d), L.ptr(daq), L.ptr(xq), None, L.ptr(bn.weight), L.ptr(bn.bias),
           L.ptr(bn.running_mean), L.ptr(bn.running_var), L.ptr(red), st)
    L.call("hg_bn_bwd_apply", C.byref(d), L.ptr(daq), L.ptr(xq), None, L.ptr(bn.weight), L.ptr(bn.bias),
           L.ptr(bn.running_mean), L.ptr(bn.running_var), L.ptr(red), None, L.ptr(dx), L.ptr(dg), L.ptr(db), None, st)
    close(nchw(dx, Cc), xr.grad, tol(dtype), "bn eval dx")
    close(dg, bn.weight.grad, 1e-4, "bn eval dgamma")
    close(db, bn.bias.grad, 1e-4, "bn eval dbeta")


def test_batchnorm_eval_and_running_update():
    torch.manual_seed(0)
    dev = "cuda"
    N, Cc, H, W = 2, 128, 16, 16
    bn = torch.nn.BatchNorm2d(Cc).to(dev)
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2.0)
    ref_bn = torch.nn.BatchNorm2d(Cc).to(dev)
    ref_bn.load_state_dict(bn.state_dict())
    xs = [torch.randn(N, Cc, H, W, device=dev) * (i + 1) for i in range(3)]
    st = L.stream_ptr()
    d = L.HgBnDesc(N * H * W, Cc, L.HG_F32, 1e-5, 1, 1)
    xq = nhwc(xs[0], torch.float32)
    y = torch.empty_like(xq)
    L.call("hg_bn_apply", C.byref(d), L.ptr(xq), None, L.ptr(bn.weight), L.ptr(bn.bias), L.ptr(bn.running_mean),
           L.ptr(bn.running_var), L.ptr(y), st)
    ref_bn.eval()
    close(nchw(y, Cc), F.relu(ref_bn(xs[0])).detach(), 1e-5, "bn eval fwd")
    # three call sites of one module, applied in order by one launch
    ref_bn.train()
    stats = []
    dtr = L.HgBnDesc(N * H * W, Cc, L.HG_F32, 1e-5, 1, 0)
    for x in xs:
        ref_bn(x)
        s = torch.zeros(3 * Cc, device=dev)
        L.call("hg_bn_stats", C.byref(dtr), L.ptr(nhwc(x, torch.float32)), L.ptr(s), st)
        stats.append(s)
    sites = (L.HgBnRunningSite * 3)(*[L.HgBnRunningSite(s.data_ptr(), float(N * H * W), 0) for s in stats])
    mods = (L.HgBnRunningModule * 1)(L.HgBnRunningModule(bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                                                         bn.num_batches_tracked.data_ptr(), Cc, Cc, 0, 3, 0.1, 0))
    sd = torch.frombuffer(bytearray(bytes(sites)), dtype=torch.uint8).cuda()
    md = torch.frombuffer(bytearray(bytes(mods)), dtype=torch.uint8).cuda()
    L.call("hg_bn_update_running", L.ptr(md), L.ptr(sd), 1, st)
    close(bn.running_mean, ref_bn.running_mean, 1e-5, "running_mean")
    close(bn.running_var, ref_bn.running_var, 1e-5, "running_var")
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked) == 3


@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool_fwd_bwd_first_max_tiebreak(dtype):
    torch.manual_seed(0)
    dev = "cuda"
    N, Cc, H, W = 2, 64, 8, 8
    x = torch.randint(0, 3, (N, Cc, H, W), device=dev).float()  # many ties
    xq = nhwc(x, dtype)
    y = torch.empty(N, H // 2, W // 2, 64, device=dev, dtype=dtype)
    st = L.stream_ptr()
    pstats = torch.zeros(3 * 64, device=dev)
    L.call("hg_maxpool2_fwd", L.hg_dtype(dtype), L.ptr(xq), N, H, W, Cc, L.ptr(y), L.ptr(pstats), st)
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 2)
    assert torch.equal(nchw(y, Cc), ref.detach())
    # fused BatchNorm statistics of the pooled tensor
    close(pstats[:Cc], ref.detach().sum((0, 2, 3)), 1e-5, "pool stats sum")
    close(pstats[64:64 + Cc], (ref.detach() ** 2).sum((0, 2, 3)), 1e-5, "pool stats sumsq")
    dy = torch.randn(N, Cc, H // 2, W // 2, device=dev)
    dyq = nhwc(dy, dtype)
    ref.backward(nchw(dyq, Cc))
    addq = nhwc(torch.randn(N, Cc, H, W, device=dev), dtype)
    dx = torch.empty_like(xq)
    L.call("hg_maxpool2_bwd", L.hg_dtype(dtype), L.ptr(xq), L.ptr(dyq), L.ptr(addq), N, H, W, Cc, L.ptr(dx), st)
    close(nchw(dx, Cc), xr.grad + nchw(addq, Cc), tol(dtype), "maxpool bwd")
    dx0 = torch.empty_like(xq)
    L.call("hg_maxpool2_bwd", L.hg_dtype(dtype), L.ptr(xq), L.ptr(dyq), None, N, H, W, Cc, L.ptr(dx0), st)
    assert torch.equal(nchw(dx0, Cc), xr.grad), "gradient must go to the first row-major maximum"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("hw,Cc", [(4, 64), (8, 64), (32, 64), (32, 256), (16, 128), (5, 96)])
@pytest.mark.parametrize("sep", [1, 0])
def test_upsample_add_fwd_bwd(dtype, mode, hw, Cc, sep):
    """x2 up-sampling (+ skip add, + statistics) and its adjoint (+ addend) against F.interpolate / autograd, through the
    separable shared-memory kernels (sep = 1, the default) and the gather kernels (sep = 0)."""
    torch.manual_seed(0)
    dev = "cuda"
    N = 2
    Cp = L.pad64(Cc)
    low = torch.randn(N, Cc, hw, hw, device=dev)
    skip = torch.randn(N, Cc, 2 * hw, 2 * hw, device=dev)
    lq, sq = nhwc(low, dtype), nhwc(skip, dtype)
    out = torch.empty_like(sq)
    st = L.stream_ptr()
    ustats = torch.zeros(3 * Cp, device=dev)
    L.call("hg_set_option", b"upsample_sep", sep)
    try:
        L.call("hg_upsample2x_add_fwd", L.hg_dtype(dtype), mode, L.ptr(lq), L.ptr(sq), N, hw, hw, Cc, L.ptr(out),
               L.ptr(ustats), st)
        dout = nhwc(torch.randn(N, Cc, 2 * hw, 2 * hw, device=dev), dtype)
        dlow = torch.empty_like(lq)
        L.call("hg_upsample2x_bwd", L.hg_dtype(dtype), mode, L.ptr(dout), None, N, hw, hw, Cc, L.ptr(dlow), st)
        addq = nhwc(torch.randn(N, Cc, hw, hw, device=dev), dtype)
        dlow2 = torch.empty_like(lq)
        L.call("hg_upsample2x_bwd", L.hg_dtype(dtype), mode, L.ptr(dout), L.ptr(addq), N, hw, hw, Cc, L.ptr(dlow2), st)
    finally:
        L.call("hg_set_option", b"upsample_sep", 1)
    ov = nchw(out, Cc)
    close(ustats[:Cc], ov.sum((0, 2, 3)), 1e-4, "upsample stats sum (of the stored values)")
    close(ustats[Cp:Cp + Cc], (ov * ov).sum((0, 2, 3)), 1e-4, "upsample stats sumsq")
    lr = nchw(lq, Cc).requires_grad_(True)
    if mode == 0:
        up = F.interpolate(lr, scale_factor=2, mode="bilinear", align_corners=True)
    else:
        up = F.interpolate(lr, scale_factor=2, mode="nearest")
    ref = up + nchw(sq, Cc)
    close(nchw(out, Cc), ref.detach(), tol(dtype), "upsample+add fwd")
    up.backward(nchw(dout, Cc))
    close(nchw(dlow, Cc), lr.grad, tol(dtype) if dtype == torch.bfloat16 else 1e-5, "upsample bwd")
    close(nchw(dlow2, Cc), lr.grad + nchw(addq, Cc), tol(dtype) if dtype == torch.bfloat16 else 1e-5, "upsample bwd + addend")


@pytest.mark.parametrize("dtype", DTYPES)
def test_stem_fwd_bwd(dtype):
    torch.manual_seed(0)
    dev = "cuda"
    N, H, W = 2, 64, 96
    x = torch.randn(N, 3, H, W, device=dev)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3).to(dev)
    y = torch.empty(N, H // 2, W // 2, 64, device=dev, dtype=dtype)
    st = L.stream_ptr()
    L.call("hg_stem_fwd", L.hg_dtype(dtype), L.ptr(x), L.ptr(conv.weight), L.ptr(conv.bias), N, H, W, 1, L.ptr(y), st)
    # bf16 path: the tensor-core stem reads the image and the weights as bf16 (what autocast does to this convolution):
    # the reference is taken on bf16-representable operands, as for every other convolution
    xr = x.to(dtype).float()
    ref = F.relu(F.conv2d(xr, conv.weight.to(dtype).float(), conv.bias, 2, 3))
    close(nchw(y, 64), ref.detach(), 1e-2 if dtype == torch.bfloat16 else 1e-5, "stem fwd")
    if dtype == torch.bfloat16:
        close(nchw(y, 64), F.relu(conv(x)).detach(), 2e-2, "stem fwd vs fp32 operands")
    dy = nhwc(torch.randn(N, 64, H // 2, W // 2, device=dev), dtype)
    # reference backward with the kernel's own stored activation as ReLU mask
    mask = (nchw(y, 64) > 0).float()
    g = nchw(dy, 64) * mask
    ref_dw = torch.nn.grad.conv2d_weight(xr, conv.weight.shape, g, 2, 3)
    dw = torch.zeros_like(conv.weight)
    db = torch.zeros(64, device=dev)
    L.call("hg_stem_bwd", L.hg_dtype(dtype), L.ptr(x), L.ptr(y), L.ptr(dy), N, H, W, 1, L.ptr(dw), L.ptr(db), st)
    close(dw, ref_dw, 1e-4, "stem dw")
    close(db, g.sum((0, 2, 3)), 1e-4, "stem db")
    if dtype == torch.bfloat16:
        # the CUDA-core kernels (fp32 image and weights) stay selectable
        L.call("hg_set_option", b"stem_tc", 0)
        try:
            y0 = torch.empty_like(y)
            L.call("hg_stem_fwd", L.hg_dtype(dtype), L.ptr(x), L.ptr(conv.weight), L.ptr(conv.bias), N, H, W, 1, L.ptr(y0), st)
            close(nchw(y0, 64), F.relu(conv(x)).detach(), 1e-2, "stem fwd (CUDA cores)")
            dw0, db0 = torch.zeros_like(conv.weight), torch.zeros(64, device=dev)
            L.call("hg_stem_bwd", L.hg_dtype(dtype), L.ptr(x), L.ptr(y), L.ptr(dy), N, H, W, 1, L.ptr(dw0), L.ptr(db0), st)
            close(dw0, torch.nn.grad.conv2d_weight(x, conv.weight.shape, g, 2, 3), 1e-4, "stem dw (CUDA cores)")
        finally:
            L.call("hg_set_option", b"stem_tc", 1)


def test_layout_roundtrip_and_add():
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(2, 17, 8, 8, device=dev)
    st = L.stream_ptr()
    for dtype in DTYPES:
        q = torch.full((2, 8, 8, 64), float("nan"), device=dev, dtype=dtype)
        L.call("hg_nchw_f32_to_nhwc", L.hg_dtype(dtype), L.ptr(x), None, 2, 17, 8, 8, L.ptr(q), st)
        assert torch.equal(q, nhwc(x, dtype))
        back = torch.empty_like(x)
        L.call("hg_nhwc_to_nchw_f32", L.hg_dtype(dtype), L.ptr(q), 2, 17, 8, 8, L.ptr(back), st)
        assert torch.equal(back, x.to(dtype).float())
        s = torch.empty_like(q)
        L.call("hg_add", L.hg_dtype(dtype), L.ptr(q), L.ptr(q), L.ptr(s), C.c_longlong(q.numel()), st)
        assert torch.equal(s.float(), (q.float() * 2).to(dtype).float())


def test_conv_with_output_batchnorm_eval_persistent_kernel():
    """The same inference epilogue through conv_persist_kernel (forced for every size): pixel-major (1x1, Np = 64 / 128 /
    256) and transposed (3x3, 128 output channels)."""
    _persist(1, p3=1)
    try:
        n0 = L.load().hg_launch_count()
        for c in [(2, 32, 32, 128, 128, 3, True), (2, 32, 32, 256, 128, 1, True), (3, 16, 16, 128, 256, 1, False),
                  (4, 16, 16, 64, 64, 1, True), (19, 64, 64, 128, 128, 3, True)]:
            test_conv_with_output_batchnorm_eval(c)
        assert L.load().hg_launch_count() > n0
    finally:
        _persist(512)


@pytest.mark.parametrize("case", [
    # N, H, W, Cin, Cout, k, relu
    (4, 64, 64, 256, 128, 1, True), (4, 32, 32, 128, 128, 3, True), (32, 4, 4, 128, 256, 1, False),
    (2, 16, 16, 256, 256, 1, True), (3, 8, 8, 64, 38, 1, True),
])
def test_conv_with_output_batchnorm_eval(case):
    """hg_conv_fprop_bnout (conv -> eval-mode BN -> ReLU in one kernel, the inference form of
    try_with_torch.py:196-205,243-256) against PyTorch fp32 ops on the same bf16-representable inputs and against
    the un-fused kernels of the same library."""
    N, H, W, Cin, Cout, k, relu = case
    torch.manual_seed(5)
    dev, dtype = "cuda", torch.bfloat16
    pad = k // 2
    x = torch.randn(N, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device=dev)
    gamma = torch.rand(Cout, device=dev) + 0.5
    beta = torch.randn(Cout, device=dev) * 0.3
    rmean = torch.randn(Cout, device=dev) * 0.2
    rvar = torch.rand(Cout, device=dev) * 2 + 0.5
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 1, pad, 1, L.HG_BF16)
    assert L.load().hg_conv_tc_eligible(C.byref(d)) == 1
    Cin_p, Cout_p = L.pad64(Cin), L.pad64(Cout)
    xq = nhwc(x, dtype)
    wf = torch.empty(k * k, Cout_p, Cin_p, device=dev, dtype=dtype)
    bias_p = torch.zeros(Cout_p, device=dev)
    bias_p[:Cout] = b
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), None, st)
    fold = L.HgBnFold(None, gamma.data_ptr(), beta.data_ptr(), rmean.data_ptr(), rvar.data_ptr(), 1e-5,
                      1 if relu else 0, 1, 0)
    y1 = torch.full((N, H, W, Cout_p), float("nan"), device=dev, dtype=dtype)
    L.call("hg_conv_fprop_bnout", C.byref(d), C.byref(fold), L.ptr(xq), L.ptr(wf), L.ptr(bias_p), L.ptr(y1), None, st)
    ref = F.batch_norm(F.conv2d(nchw(xq, Cin), w.to(dtype).float(), b, 1, pad), rmean, rvar, gamma, beta, False, 0.0,
                       1e-5)
    if relu:
        ref = F.relu(ref)
    close(nchw(y1, Cout), ref, 2e-2, "conv + output BN vs torch")
    assert (y1[..., Cout:] == 0).all()  # padded channels stay zero
    # un-fused: conv (bf16 store) -> hg_bn_apply; the fused kernel skips one bf16 rounding, so only close
    y0 = torch.empty_like(y1)
    L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xq), L.ptr(wf), L.ptr(bias_p), None, L.ptr(y0), None, None, st)
    bnd = L.HgBnDesc(N * H * W, Cout, L.HG_BF16, 1e-5, 1 if relu else 0, 1)
    a0 = torch.empty_like(y0)
    L.call("hg_bn_apply", C.byref(bnd), L.ptr(y0), None, L.ptr(gamma), L.ptr(beta), L.ptr(rmean), L.ptr(rvar),
           L.ptr(a0), st)
    close(y1.float(), a0.float(), 2e-2, "fused vs un-fused")
    # a training-mode BatchNorm cannot be folded into its producer: loud error, no fallback
    bad = L.HgBnFold(None, gamma.data_ptr(), beta.data_ptr(), rmean.data_ptr(), rvar.data_ptr(), 1e-5, 1, 0, 0)
    with pytest.raises(RuntimeError, match="inference-mode"):
        L.call("hg_conv_fprop_bnout", C.byref(d), C.byref(bad), L.ptr(xq), L.ptr(wf), L.ptr(bias_p), L.ptr(y1), None, st)
