"""Do the large-map streaming kernels run faster when their operands are L2-resident?  Each kernel is timed on ONE
buffer set (input + output <= 126 MB stay in L2 across launches) and on rotating sets larger than L2 (every launch
streams from HBM).  If both times agree, the kernel is not bound by HBM but by its own latency / occupancy."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

B, DEV, BF = 32, "cuda", torch.bfloat16
REPS = 20


def timeit(fn, nrot):
    for i in range(3):
        fn(i % nrot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):
        fn(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / REPS


def bn(H, Cc, NR):
    M = B * H * H
    xs = [torch.randn(M, Cc, device=DEV, dtype=BF) for _ in range(NR)]
    ys = [torch.empty(M, Cc, device=DEV, dtype=BF) for _ in range(NR)]
    gs = [torch.randn(M, Cc, device=DEV, dtype=BF) for _ in range(NR)]
    gamma, beta = torch.ones(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    stats, red = torch.zeros(3 * Cc, device=DEV), torch.zeros(2 * Cc, device=DEV)
    dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    d = L.HgBnDesc(M, Cc, L.HG_BF16, 1e-5, 1, 0)
    st = L.stream_ptr()
    L.call("hg_bn_stats", C.byref(d), L.ptr(xs[0]), L.ptr(stats), st)
    L.call("hg_bn_bwd_reduce", C.byref(d), L.ptr(gs[0]), L.ptr(xs[0]), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None,
           L.ptr(red), st)
    f1 = lambda i: L.call("hg_bn_apply", C.byref(d), L.ptr(xs[i]), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None,
                          L.ptr(ys[i]), st)
    f2 = lambda i: L.call("hg_bn_bwd_apply", C.byref(d), L.ptr(gs[i]), L.ptr(xs[i]), L.ptr(stats), L.ptr(gamma),
                          L.ptr(beta), None, None, L.ptr(red), None, L.ptr(ys[i]), L.ptr(dg), L.ptr(db), None, st)
    for name, f, nb in (("bn_apply", f1, 2), ("bn_bwd_apply", f2, 3)):
        mb = nb * M * Cc * 2 / 1e6
        t1, tn = timeit(f, 1), timeit(f, NR)
        print(f"{name:14s} C{Cc} @{H}: {mb:6.1f} MB  L2-resident {t1:7.2f} us ({mb / t1 * 1e-3:5.2f} TB/s)   rotating {tn:7.2f} us "
              f"({mb / tn * 1e-3:5.2f} TB/s)", flush=True)


def conv(H, Cin, Cout, k, NR):
    d = L.HgConvDesc(B, H, H, Cin, Cout, k, k, 1, k // 2, 1, L.HG_BF16)
    xs = [torch.randn(B, H, H, Cin, device=DEV).to(BF) for _ in range(NR)]
    ys = [torch.empty(B, H, H, Cout, device=DEV, dtype=BF) for _ in range(NR)]
    wf = (torch.randn(k * k, Cout, Cin, device=DEV) * 0.05).to(BF)
    bias = torch.zeros(Cout, device=DEV)
    stats = torch.zeros(3 * Cout, device=DEV)
    st = L.stream_ptr()
    f = lambda i: L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xs[i]), L.ptr(wf), L.ptr(bias), None, L.ptr(ys[i]),
                         L.ptr(stats), None, st)
    mb = B * H * H * (Cin + Cout) * 2 / 1e6
    t1, tn = timeit(f, 1), timeit(f, NR)
    print(f"fprop {Cin}->{Cout} k{k} @{H}: {mb:6.1f} MB  L2-resident {t1:7.2f} us ({mb / t1 * 1e-3:5.2f} TB/s)   rotating {tn:7.2f} us "
          f"({mb / tn * 1e-3:5.2f} TB/s)", flush=True)


if __name__ == "__main__":
    L.call("hg_set_option", b"persist_3x3", 0)
    print(torch.cuda.get_device_name(0))
    bn(64, 128, 8)
    bn(64, 256, 6)
    bn(32, 128, 16)
    bn(32, 256, 12)
    conv(64, 256, 128, 1, 6)
    conv(64, 128, 128, 3, 8)
    conv(32, 256, 128, 1, 12)
    conv(32, 128, 128, 3, 12)
    # copy bandwidth reference: L2-resident vs HBM
    for mb in (32, 64, 256):
        n = mb * 1024 * 1024 // 2
        a = [torch.empty(n, device=DEV, dtype=BF) for _ in range(6)]
        b = [torch.empty(n, device=DEV, dtype=BF) for _ in range(6)]
        f = lambda i: b[i].copy_(a[i])
        t1, tn = timeit(f, 1), timeit(f, 6)
        print(f"torch copy {mb} MiB -> {mb} MiB: L2-resident {t1:7.2f} us ({2 * mb * 1.048576 / t1 * 1e-3:5.2f} TB/s)  rotating {tn:7.2f} us "
              f"({2 * mb * 1.048576 / tn * 1e-3:5.2f} TB/s)")
