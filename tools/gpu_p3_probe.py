"""Where conv3x3_persist_kernel spends its time (DBG build: make -C .../csrc DBG=1): per-role barrier-wait and phase
cycles of CTA 0, and the launch time with the epilogue / the MMAs switched off (p3_dbg 1 / 2 / 3)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

B = int(os.environ.get("B", "32"))
DEV, BF = "cuda", torch.bfloat16


def run(H, Cc, kind, dbg):
    d = L.HgConvDesc(B, H, H, Cc, Cc, 3, 3, 1, 1, 1, L.HG_BF16)
    M = B * H * H
    NROT = 4
    xs = [torch.randn(B, H, H, Cc, device=DEV).to(BF) for _ in range(NROT)]
    ys = [torch.randn(B, H, H, Cc, device=DEV).to(BF) for _ in range(NROT)]
    gs = [torch.empty(B, H, H, Cc, device=DEV, dtype=BF) for _ in range(NROT)]
    wf = (torch.randn(9, Cc, Cc, device=DEV) * 0.05).to(BF)
    bias = torch.zeros(Cc, device=DEV)
    stats = torch.zeros(3 * Cc, device=DEV)
    xstats = torch.zeros(3 * Cc, device=DEV)
    bnd = L.HgBnDesc(M, Cc, L.HG_BF16, 1e-5, 1, 0)
    st = L.stream_ptr()
    L.call("hg_bn_stats", C.byref(bnd), L.ptr(xs[0]), L.ptr(xstats), st)
    gam, bet = torch.ones(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    fold = L.HgBnFold(xstats.data_ptr(), gam.data_ptr(), bet.data_ptr(), None, None, 1e-5, 1, 0, 0)
    red = torch.zeros(2 * Cc, device=DEV)
    if kind == "fprop":
        fn = lambda i: L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xs[i]), L.ptr(wf), L.ptr(bias), None, L.ptr(ys[i]),
                              L.ptr(stats), None, st)
    else:
        fn = lambda i: L.call("hg_conv_dgrad_bn", C.byref(d), C.byref(fold), L.ptr(ys[i]), L.ptr(wf), L.ptr(xs[i]),
                              L.ptr(gs[i]), L.ptr(red), st)
    L.call("hg_set_option", b"p3_dbg", dbg)
    L.call("hg_set_option", b"dbg_ts", 1)
    fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for i in range(reps):
        fn((i + 1) % NROT)
    e1.record()
    torch.cuda.synchronize()
    print(f"{kind} {Cc}->{Cc} @{H} p3_dbg={dbg}: {e0.elapsed_time(e1) / reps * 1e3:8.2f} us", flush=True)
    sys.stdout.flush()
    L.call("hg_set_option", b"dbg_ts", 3)
    L.call("hg_set_option", b"p3_dbg", 0)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), f"B={B}")
    for H in (64, 32):
        for kind in ("fprop", "dgrad_bn"):
            for dbg in (0, 1, 2, 3):
                run(H, 128, kind, dbg)
