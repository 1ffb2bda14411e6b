"""Where conv_persist_kernel spends its time (DBG build: make -C .../csrc DBG=1): per-role barrier-wait and phase
cycles of CTA 0, and the launch time with the epilogue / the MMAs switched off (ps_dbg 1 / 2 / 3)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

B = int(os.environ.get("B", "32"))
DEV, BF = "cuda", torch.bfloat16


def run(H, Cin, Cout, k, res, kind, dbg):
    d = L.HgConvDesc(B, H, H, Cin, Cout, k, k, 1, k // 2, 1, L.HG_BF16)
    M = B * H * H
    NROT = 4
    xs = [torch.randn(B, H, H, Cin, device=DEV).to(BF) for _ in range(NROT)]
    ys = [torch.randn(B, H, H, Cout, device=DEV).to(BF) for _ in range(NROT)]
    rs = [torch.randn(B, H, H, Cout, device=DEV).to(BF) for _ in range(NROT)]
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
    if kind == "fprop":
        fn = lambda i: L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(xs[i]), L.ptr(wf), L.ptr(bias),
                              L.ptr(rs[i]) if res else None, L.ptr(ys[i]), L.ptr(stats), None, st)
    else:
        fn = lambda i: L.call("hg_conv_dgrad_bn", C.byref(d), C.byref(fold), L.ptr(ys[i]), L.ptr(wd), L.ptr(xs[i]),
                              L.ptr(gs[i]), L.ptr(red), st)
    L.call("hg_set_option", b"ps_dbg", dbg)
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
    print(f"{kind} {Cin}->{Cout} k{k} res={int(res)} @{H} ps_dbg={dbg}: {e0.elapsed_time(e1) / reps * 1e3:8.2f} us", flush=True)
    sys.stdout.flush()
    L.call("hg_set_option", b"dbg_ts", 3)
    L.call("hg_set_option", b"ps_dbg", 0)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), f"B={B}")
    L.call("hg_set_option", b"persist_3x3", 1)
    shapes = os.environ.get("SHAPES", "64:256:128:1:0,64:128:256:1:1,64:256:256:1:0,32:256:128:1:0,32:128:256:1:1")
    dbgs = [int(v) for v in os.environ.get("DBGS", "0,1,2,3").split(",")]
    for s in shapes.split(","):
        H, ci, co, k, r = (int(v) for v in s.split(":"))
        for kind in ("fprop", "dgrad_bn"):
            for dbg in dbgs:
                run(H, ci, co, k, bool(r), kind, dbg)
