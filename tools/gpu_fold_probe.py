"""Plain vs BatchNorm-folded convolution kernels on the two dominant shapes (timing with CUDA events; also the
command ncu captures)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda")
DT = torch.bfloat16
REPS = int(os.environ.get("REPS", "10"))


def timeit(fn, reps=REPS):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def run(B, hw, cin, cout, k):
    d = L.HgConvDesc(B, hw, hw, cin, cout, k, k, 1, k // 2, 1, L.HG_BF16)
    M = B * hw * hw
    st = L.stream_ptr()
    # rotate over several buffers so the working set exceeds L2 at 64x64
    nbuf = 6 if hw >= 32 else 1
    xs = [torch.randn(B, hw, hw, cin, device=dev).to(DT) for _ in range(nbuf)]
    ys = [torch.zeros(B, hw, hw, cout, device=dev, dtype=DT) for _ in range(nbuf)]
    dxs = [torch.zeros(B, hw, hw, cin, device=dev, dtype=DT) for _ in range(nbuf)]
    wf = (torch.randn(k * k, cout, cin, device=dev) * 0.05).to(DT)
    wd = (torch.randn(k * k, cin, cout, device=dev) * 0.05).to(DT)
    bias = torch.zeros(cout, device=dev)
    stats = torch.zeros(3 * cout, device=dev)
    xstats = torch.zeros(3 * cin, device=dev)
    bnd = L.HgBnDesc(M, cin, L.HG_BF16, 1e-5, 1, 0)
    L.call("hg_bn_stats", C.byref(bnd), L.ptr(xs[0]), L.ptr(xstats), st)
    gam, bet = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
    fold = L.HgBnFold(xstats.data_ptr(), gam.data_ptr(), bet.data_ptr(), None, None, 1e-5, 1, 0, 0)
    red = torch.zeros(2 * cin, device=dev)
    dw = torch.zeros(k * k, cout, cin, device=dev)
    it = [0]

    def nxt():
        it[0] = (it[0] + 1) % nbuf
        return it[0]

    res = {}
    res["fprop"] = timeit(lambda: L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xs[nxt()]), L.ptr(wf), L.ptr(bias), None,
                                         L.ptr(ys[it[0]]), L.ptr(stats), None, st))
    res["fprop_bn"] = timeit(lambda: L.call("hg_conv_fprop_bn", C.byref(d), C.byref(fold), L.ptr(xs[nxt()]), L.ptr(wf),
                                            L.ptr(bias), None, L.ptr(ys[it[0]]), L.ptr(stats), None, st))
    res["dgrad"] = timeit(lambda: L.call("hg_conv_dgrad", C.byref(d), L.ptr(ys[nxt()]), L.ptr(wd), None,
                                         L.ptr(dxs[it[0]]), st))
    res["dgrad_bn"] = timeit(lambda: L.call("hg_conv_dgrad_bn", C.byref(d), C.byref(fold), L.ptr(ys[nxt()]), L.ptr(wd),
                                            L.ptr(xs[it[0]]), L.ptr(dxs[it[0]]), L.ptr(red), st))
    res["wgrad"] = timeit(lambda: L.call("hg_conv_wgrad", C.byref(d), L.ptr(xs[nxt()]), L.ptr(ys[it[0]]), L.ptr(dw),
                                         None, st))
    res["wgrad_bn"] = timeit(lambda: L.call("hg_conv_wgrad_bn", C.byref(d), C.byref(fold), L.ptr(xs[nxt()]),
                                            L.ptr(ys[it[0]]), L.ptr(dw), None, st))
    print(f"B={B} {hw}x{hw} {cin}->{cout} k{k}: " + "  ".join(f"{n} {t:7.1f}us" for n, t in res.items()), flush=True)


if __name__ == "__main__":
    shapes = [(32, 64, 128, 128, 3), (32, 64, 256, 128, 1), (32, 64, 128, 256, 1)]
    if os.environ.get("ALL"):
        shapes += [(32, 32, 128, 128, 3), (32, 16, 128, 128, 3), (32, 4, 128, 128, 3), (32, 4, 256, 128, 1)]
    for s in shapes:
        run(*s)
