"""The large-map kernels that bound the training step (64x64 / 32x32 levels, B=32), launched a few times each on
rotating buffers (> L2) through the C ABI: the command `ncu --set full` captures for profiles/r02_top_kernels_*.
SHAPES env: comma list of H:Cin:Cout:k:res (default: the six dominant shapes)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

B = int(os.environ.get("B", "32"))
REPS = int(os.environ.get("REPS", "3"))
NROT = 6
DEV, BF = "cuda", torch.bfloat16
KINDS = os.environ.get("KINDS", "fprop,dgrad_bn,wgrad").split(",")


def run(H, Cin, Cout, k, res):
    d = L.HgConvDesc(B, H, H, Cin, Cout, k, k, 1, k // 2, 1, L.HG_BF16)
    M = B * H * H
    xs = [torch.randn(B, H, H, Cin, device=DEV).to(BF) for _ in range(NROT)]
    ys = [torch.randn(B, H, H, Cout, device=DEV).to(BF) for _ in range(NROT)]
    rs = [torch.randn(B, H, H, Cout, device=DEV).to(BF) for _ in range(NROT)] if res else None
    gs = [torch.empty(B, H, H, Cin, device=DEV, dtype=BF) for _ in range(NROT)]
    wf = (torch.randn(k * k, Cout, Cin, device=DEV) * 0.05).to(BF)
    wd = (torch.randn(k * k, Cin, Cout, device=DEV) * 0.05).to(BF)
    bias = torch.zeros(Cout, device=DEV)
    stats = torch.zeros(3 * Cout, device=DEV)
    xstats = torch.zeros(3 * Cin, device=DEV)
    bnd = L.HgBnDesc(M, Cin, L.HG_BF16, 1e-5, 1, 0)
    st = L.stream_ptr()
    L.call("hg_bn_stats", C.byref(bnd), L.ptr(xs[0]), L.ptr(xstats), st)
    gam, bet = torch.ones(Cin, device=DEV), torch.zeros(Cin, device=DEV)
    fold = L.HgBnFold(xstats.data_ptr(), gam.data_ptr(), bet.data_ptr(), None, None, 1e-5, 1, 0, 0)
    red = torch.zeros(2 * Cin, device=DEV)
    dw = torch.zeros(k * k, Cout, Cin, device=DEV)
    calls = {
        "fprop": lambda i: L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xs[i]), L.ptr(wf), L.ptr(bias),
                                  L.ptr(rs[i]) if res else None, L.ptr(ys[i]), L.ptr(stats), None, st),
        "dgrad_bn": lambda i: L.call("hg_conv_dgrad_bn", C.byref(d), C.byref(fold), L.ptr(ys[i]), L.ptr(wd), L.ptr(xs[i]),
                                     L.ptr(gs[i]), L.ptr(red), st),
        "wgrad": lambda i: L.call("hg_conv_wgrad", C.byref(d), L.ptr(xs[i]), L.ptr(ys[i]), L.ptr(dw), None, st),
    }
    for kind in KINDS:
        fn = calls[kind]
        fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(REPS):
            fn((i + 1) % NROT)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / REPS * 1e3
        fl = 2.0 * M * Cin * Cout * k * k
        print(f"{kind:9s} {Cin}->{Cout} k{k} @{H} res={int(res)}: {us:8.2f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)


def main():
    for opt in os.environ.get("HG_OPTIONS", "").split(","):   # e.g. HG_OPTIONS=wgrad_kpx=128
        if "=" in opt:
            k_, v_ = opt.split("=")
            L.call("hg_set_option", k_.encode(), int(v_))
    shapes = os.environ.get("SHAPES", "64:128:128:3:0,64:128:256:1:1,64:256:128:1:0,64:256:256:1:0,32:128:128:3:0,32:128:256:1:1")
    print(torch.cuda.get_device_name(0), f"B={B}")
    for s in shapes.split(","):
        H, ci, co, k, r = (int(v) for v in s.split(":"))
        run(H, ci, co, k, bool(r))


if __name__ == "__main__":
    main()
