"""Per-kernel micro-benchmark through the C ABI (CUDA events on the launching stream, rotating buffers larger
than L2 so every launch reads from HBM).  Prints achieved TFLOP/s and GB/s per shape of the 8-stack step."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

B = 32
NROT = 6  # rotating copies of the big tensors (6 x 67 MB >> 126 MB L2)


def timeit(fn, iters=12, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def conv_case(H, Cin, Cout, k, res):
    dev = "cuda"
    d = L.HgConvDesc(B, H, H, Cin, Cout, k, k, 1, k // 2, 1, L.HG_BF16)
    cin_p, cout_p = L.pad64(Cin), L.pad64(Cout)
    M = B * H * H
    nrot = NROT if M * max(cin_p, cout_p) * 2 > 8e6 else 1
    xs = [torch.randn(B, H, H, cin_p, device=dev, dtype=torch.bfloat16) for _ in range(nrot)]
    ys = [torch.randn(B, H, H, cout_p, device=dev, dtype=torch.bfloat16) for _ in range(nrot)]
    rs = [torch.randn(B, H, H, cout_p, device=dev, dtype=torch.bfloat16) for _ in range(nrot)] if res else None
    w = torch.randn(Cout, Cin, k, k, device=dev) * 0.05
    wf = torch.empty(k * k, cout_p, cin_p, device=dev, dtype=torch.bfloat16)
    wd = torch.empty(k * k, cin_p, cout_p, device=dev, dtype=torch.bfloat16)
    bias = torch.zeros(cout_p, device=dev)
    stats = torch.zeros(3 * cout_p, device=dev)
    dwp = torch.zeros(k * k, cout_p, cin_p, device=dev)
    dbias = torch.zeros(cout_p, device=dev)
    st = L.stream_ptr()
    L.call("hg_pack_conv_weight", C.byref(d), L.ptr(w), L.ptr(wf), L.ptr(wd), st)
    flops = 2.0 * M * Cin * Cout * k * k
    out = {}
    t = timeit(lambda i: L.call("hg_conv_fprop", C.byref(d), L.ptr(xs[i % nrot]), L.ptr(wf), L.ptr(bias),
                                L.ptr(rs[i % nrot]) if res else None, L.ptr(ys[i % nrot]), L.ptr(stats), st))
    byt = M * (cin_p + cout_p * (2 if res else 1)) * 2
    out["fprop"] = (t, flops / t / 1e6, byt / t / 1e3)
    t = timeit(lambda i: L.call("hg_conv_dgrad", C.byref(d), L.ptr(ys[i % nrot]), L.ptr(wd), None, L.ptr(xs[i % nrot]), st))
    out["dgrad"] = (t, flops / t / 1e6, M * (cin_p + cout_p) * 2 / t / 1e3)
    t = timeit(lambda i: L.call("hg_conv_wgrad", C.byref(d), L.ptr(xs[i % nrot]), L.ptr(ys[i % nrot]), L.ptr(dwp), None, st))
    out["wgrad"] = (t, flops / t / 1e6, M * (cin_p + cout_p) * 2 / t / 1e3)
    t = timeit(lambda i: L.call("hg_conv_wgrad", C.byref(d), L.ptr(xs[i % nrot]), L.ptr(ys[i % nrot]), None, L.ptr(dbias), st))
    out["dbias"] = (t, 0.0, M * cout_p * 2 / t / 1e3)
    return out


def bn_case(H, Cc):
    dev = "cuda"
    M = B * H * H
    cp = L.pad64(Cc)
    nrot = NROT if M * cp * 2 > 8e6 else 1
    xs = [torch.randn(M, cp, device=dev, dtype=torch.bfloat16) for _ in range(nrot)]
    ys = [torch.empty(M, cp, device=dev, dtype=torch.bfloat16) for _ in range(nrot)]
    gs = [torch.randn(M, cp, device=dev, dtype=torch.bfloat16) for _ in range(nrot)]
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    stats = torch.zeros(3 * cp, device=dev)
    red = torch.zeros(3 * cp, device=dev)
    dg, db = torch.zeros(Cc, device=dev), torch.zeros(Cc, device=dev)
    d = L.HgBnDesc(M, Cc, L.HG_BF16, 1e-5, 1, 0)
    st = L.stream_ptr()
    L.call("hg_bn_stats", C.byref(d), L.ptr(xs[0]), L.ptr(stats), st)
    eb = M * cp * 2
    out = {}
    t = timeit(lambda i: L.call("hg_bn_stats", C.byref(d), L.ptr(xs[i % nrot]), L.ptr(red), st))
    out["stats"] = (t, 0, eb / t / 1e3)
    t = timeit(lambda i: L.call("hg_bn_apply", C.byref(d), L.ptr(xs[i % nrot]), L.ptr(stats), L.ptr(gamma), L.ptr(beta),
                                None, None, L.ptr(ys[i % nrot]), st))
    out["apply"] = (t, 0, 2 * eb / t / 1e3)
    t = timeit(lambda i: L.call("hg_bn_bwd_reduce", C.byref(d), L.ptr(gs[i % nrot]), L.ptr(xs[i % nrot]), L.ptr(stats),
                                L.ptr(gamma), L.ptr(beta), None, None, L.ptr(red), st))
    out["bwd_reduce"] = (t, 0, 2 * eb / t / 1e3)
    t = timeit(lambda i: L.call("hg_bn_bwd_apply", C.byref(d), L.ptr(gs[i % nrot]), L.ptr(xs[i % nrot]), L.ptr(stats),
                                L.ptr(gamma), L.ptr(beta), None, None, L.ptr(red), None, L.ptr(ys[i % nrot]), L.ptr(dg),
                                L.ptr(db), None, st))
    out["bwd_apply"] = (t, 0, 3 * eb / t / 1e3)
    return out


def main():
    print(torch.cuda.get_device_name(0))
    res = {}
    print(f"{'op':34s} {'us':>9s} {'TFLOP/s':>9s} {'GB/s':>9s}")
    for H in (64, 32, 16, 8, 4):
        for (ci, co, k, r) in ((128, 128, 3, False), (256, 128, 1, False), (128, 256, 1, True)):
            o = conv_case(H, ci, co, k, r)
            for kind, (t, tf, gb) in o.items():
                name = f"conv {kind} {ci}->{co} k{k} @{H}"
                res[name] = (t, tf, gb)
                print(f"{name:34s} {t:9.2f} {tf:9.1f} {gb:9.0f}")
    for (H, ci, co, k, r) in ((64, 256, 256, 1, False), (64, 256, 16, 1, False), (64, 16, 256, 1, True),
                              (128, 64, 64, 3, False), (128, 64, 128, 1, True)):
        o = conv_case(H, ci, co, k, r)
        for kind, (t, tf, gb) in o.items():
            name = f"conv {kind} {ci}->{co} k{k} @{H}"
            res[name] = (t, tf, gb)
            print(f"{name:34s} {t:9.2f} {tf:9.1f} {gb:9.0f}")
    for H in (64, 32, 16, 8, 4):
        for cc in (256, 128):
            o = bn_case(H, cc)
            for kind, (t, tf, gb) in o.items():
                name = f"bn {kind} C{cc} @{H}"
                res[name] = (t, tf, gb)
                print(f"{name:34s} {t:9.2f} {'':>9s} {gb:9.0f}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/kernel_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
