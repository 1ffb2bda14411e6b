"""Bring-up probe for the tcgen05 convolution kernels (run on a B200 through gpurun).

Compares hg_conv_fprop / hg_conv_dgrad / hg_conv_wgrad (bf16 tensor-core path) with torch.nn.functional
convolutions evaluated in fp32 on the same bf16-rounded operands.
"""
import sys
import os
import ctypes as C

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402


def nhwc_pad(t_nchw, dtype):
    n, c, h, w = t_nchw.shape
    cp = L.pad64(c)
    out = torch.zeros(n, h, w, cp, device=t_nchw.device, dtype=dtype)
    out[..., :c] = t_nchw.permute(0, 2, 3, 1).to(dtype)
    return out.contiguous()


def report(name, got, ref, tol):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    ok = err <= tol * scale
    print(f"{'OK  ' if ok else 'FAIL'} {name}: max_abs_err={err:.4e} ref_max={scale:.4e} rel={err / scale:.3e}")
    if not ok:
        bad = ((got - ref).abs() > tol * scale).nonzero()
        print("   mismatches:", bad.shape[0], "of", got.numel(), "first:", bad[:8].tolist())
    return ok


def run_case(N, H, W, Cin, Cout, k, dil=1, res=False, head=False, dtype=torch.bfloat16):
    torch.manual_seed(0)
    dev = "cuda"
    pad = dil * (k // 2)
    x = torch.randn(N, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5
    b = torch.randn(Cout, device=dev)
    r = torch.randn(N, Cout, H, W, device=dev) if res else None
    hd = L.hg_dtype(dtype)
    d = L.HgConvDesc(N, H, W, Cin, Cout, k, k, 1, pad, dil, hd)
    Cin_p, Cout_p = L.pad64(Cin), L.pad64(Cout)
    xq = nhwc_pad(x, dtype)
    rq = nhwc_pad(r, dtype) if res else None
    wf = torch.empty(k * k, Cout_p, Cin_p, device=dev, dtype=dtype)
    wd = torch.empty(k * k, Cin_p, Cout_p, device=dev, dtype=dtype)
    bias_p = torch.zeros(Cout_p, device=dev)
    bias_p[:Cout] = b
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), st)
    y = torch.full((N, H, W, Cout_p), float("nan"), device=dev, dtype=dtype)
    stats = torch.zeros(3 * Cout_p, device=dev)
    nchw = torch.full((N, Cout, H, W), float("nan"), device=dev) if head else None
    L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xq), L.ptr(wf), L.ptr(bias_p), L.ptr(rq), L.ptr(y), L.ptr(stats),
           L.ptr(nchw), st)
    torch.cuda.synchronize()
    # reference on the rounded operands
    xr = xq[..., :Cin].float().permute(0, 3, 1, 2)
    wr = w.to(dtype).float()
    ref = F.conv2d(xr, wr, b, 1, pad, dil)
    if res:
        ref = ref + rq[..., :Cout].float().permute(0, 3, 1, 2)
    tag = f"N{N} {H}x{W} {Cin}->{Cout} k{k} d{dil} res={res} head={head} {dtype}"
    tol = 1.5e-2 if dtype == torch.bfloat16 else 1e-5
    ok = report("fprop " + tag, y[..., :Cout].float().permute(0, 3, 1, 2), ref, tol)
    if Cout_p > Cout:
        ok &= report("fprop pad lanes zero " + tag, y[..., Cout:].float(), torch.zeros_like(y[..., Cout:].float()), 1.0)
    if head:
        ok &= report("fprop nchw " + tag, nchw, ref, 1e-3 if dtype == torch.bfloat16 else 1e-5)
    yv = y[..., :Cout].float()
    ok &= report("stats sum " + tag, stats[:Cout], yv.sum((0, 1, 2)), 2e-3)
    ok &= report("stats sumsq " + tag, stats[Cout_p:Cout_p + Cout], (yv * yv).sum((0, 1, 2)), 2e-3)
    # dgrad
    dy = torch.randn(N, Cout, H, W, device=dev)
    dyq = nhwc_pad(dy, dtype)
    dx = torch.full((N, H, W, Cin_p), float("nan"), device=dev, dtype=dtype)
    L.call("hg_conv_dgrad", C.byref(d), L.ptr(dyq), L.ptr(wd), None, L.ptr(dx), st)
    torch.cuda.synchronize()
    dyr = dyq[..., :Cout].float().permute(0, 3, 1, 2)
    ref_dx = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, 1, pad, dil)
    ok &= report("dgrad " + tag, dx[..., :Cin].float().permute(0, 3, 1, 2), ref_dx, tol)
    # wgrad
    dwp = torch.zeros(k * k, Cout_p, Cin_p, device=dev)
    dbias = torch.zeros(Cout, device=dev)
    L.call("hg_conv_wgrad", C.byref(d), L.ptr(xq), L.ptr(dyq), L.ptr(dwp), L.ptr(dbias), st)
    dw = torch.zeros_like(w)
    L.call("hg_unpack_conv_wgrad", C.byref(d), L.ptr(dwp), L.ptr(dw), 0, st)
    torch.cuda.synchronize()
    ref_dw = torch.nn.grad.conv2d_weight(xr, w.shape, dyr, 1, pad, dil)
    ok &= report("wgrad " + tag, dw, ref_dw, 2e-3 if dtype == torch.bfloat16 else 1e-5)
    ok &= report("dbias " + tag, dbias, dyr.sum((0, 2, 3)), 2e-3)
    return ok


def main():
    lib = L.load()
    print("device ok:", lib.hg_device_ok(), torch.cuda.get_device_name(0))
    cases = [
        dict(N=2, H=16, W=16, Cin=64, Cout=64, k=1),
        dict(N=2, H=16, W=16, Cin=256, Cout=128, k=1),
        dict(N=2, H=64, W=64, Cin=128, Cout=128, k=3),
        dict(N=2, H=64, W=64, Cin=128, Cout=256, k=1, res=True),
        dict(N=4, H=4, W=4, Cin=128, Cout=128, k=3),
        dict(N=3, H=8, W=8, Cin=128, Cout=128, k=3),
        dict(N=1, H=128, W=128, Cin=64, Cout=64, k=3),
        dict(N=2, H=32, W=32, Cin=256, Cout=256, k=1, res=True),
        dict(N=2, H=64, W=64, Cin=256, Cout=16, k=1, head=True),
        dict(N=2, H=64, W=64, Cin=16, Cout=256, k=1),
        dict(N=2, H=64, W=64, Cin=256, Cout=17, k=1, head=True),
        dict(N=2, H=4, W=4, Cin=256, Cout=256, k=3, dil=6),
    ]
    all_ok = True
    if "--ref" in sys.argv:
        L.call("hg_set_option", b"force_ref_conv", 1)
    for c in cases:
        try:
            all_ok &= run_case(**c)
        except Exception as e:  # noqa: BLE001
            print("EXC ", c, repr(e))
            all_ok = False
            if "CUDA" in repr(e) or "cuda" in repr(e):
                break
    if "--f32" in sys.argv:
        for c in cases[:6]:
            all_ok &= run_case(dtype=torch.float32, **c)
    print("ALL OK" if all_ok else "SOME FAILED", "launches:", L.launch_count())
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
