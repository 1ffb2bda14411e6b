"""In-graph latency of chains of DEPENDENT small kernels (the low-resolution hourglass levels are latency-bound:
SURVEY 2.4).  Captures N back-to-back calls of one entry point into a CUDA graph and reports us per call."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

dev = torch.device("cuda")
DT = torch.bfloat16


def graph_time(body, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            body()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def conv_chain(B, hw, cin, cout, k, n=100, with_stats=True):
    d = L.HgConvDesc(B, hw, hw, cin, cout, k, k, 1, k // 2, 1, L.HG_BF16)
    x = torch.randn(B, hw, hw, cin, device=dev).to(DT)
    y = torch.zeros(B, hw, hw, cout, device=dev, dtype=DT)
    w = torch.randn(k * k, cout, cin, device=dev).to(DT) * 0.05
    bias = torch.zeros(cout, device=dev)
    stats = torch.zeros(3 * cout, device=dev)
    # ping-pong so that every call depends on the previous one when cin == cout
    bufs = [x, y] if cin == cout else None

    def body():
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(n):
            if bufs:
                a, b = bufs[i % 2], bufs[(i + 1) % 2]
            else:
                a, b = x, y
            L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(a), L.ptr(w), L.ptr(bias), None, L.ptr(b),
                   L.ptr(stats) if with_stats else None, None, st)

    return graph_time(body) / n * 1e3


def bn_chain(B, hw, c, n=100):
    d = L.HgBnDesc(B * hw * hw, c, L.HG_BF16, 1e-5, 1, 0)
    x = torch.randn(B, hw, hw, c, device=dev).to(DT)
    y = torch.zeros_like(x)
    stats = torch.zeros(3 * c, device=dev)
    stats[:c] = 0.0
    stats[c:] = float(B * hw * hw)
    gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)

    def body():
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(n):
            a, b = (x, y) if i % 2 == 0 else (y, x)
            L.call("hg_bn_apply", C.byref(d), L.ptr(a), L.ptr(stats), L.ptr(gam), L.ptr(bet), None, None, L.ptr(b), st)

    return graph_time(body) / n * 1e3


def bn_bwd_chain(B, hw, c, n=100, with_addend=False):
    d = L.HgBnDesc(B * hw * hw, c, L.HG_BF16, 1e-5, 1, 0)
    x = torch.randn(B, hw, hw, c, device=dev).to(DT)
    g = torch.randn(B, hw, hw, c, device=dev).to(DT)
    dx = torch.zeros_like(x)
    add = torch.randn(B, hw, hw, c, device=dev).to(DT) if with_addend else None
    stats = torch.zeros(3 * c, device=dev)
    stats[c:] = float(B * hw * hw)
    red = torch.zeros(2 * c, device=dev)
    gam, bet = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    dg, db, cs = torch.zeros(c, device=dev), torch.zeros(c, device=dev), torch.zeros(c, device=dev)

    def body():
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(n):
            a, b = (g, dx) if i % 2 == 0 else (dx, g)
            L.call("hg_bn_bwd_apply", C.byref(d), L.ptr(a), L.ptr(x), L.ptr(stats), L.ptr(gam), L.ptr(bet), None, None,
                   L.ptr(red), L.ptr(add), L.ptr(b), L.ptr(dg), L.ptr(db), L.ptr(cs), st)

    return graph_time(body) / n * 1e3


def wgrad_chain(B, hw, cin, cout, k, n=50):
    d = L.HgConvDesc(B, hw, hw, cin, cout, k, k, 1, k // 2, 1, L.HG_BF16)
    x = torch.randn(B, hw, hw, cin, device=dev).to(DT)
    dy = torch.randn(B, hw, hw, cout, device=dev).to(DT)
    dw = torch.zeros(k * k, cout, cin, device=dev)

    def body():
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for i in range(n):
            L.call("hg_conv_wgrad", C.byref(d), L.ptr(x), L.ptr(dy), L.ptr(dw), None, st)

    return graph_time(body) / n * 1e3


def empty_chain(n=200):
    t = torch.zeros(64, device=dev, dtype=DT)

    def body():
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(n):
            L.call("hg_add", L.HG_BF16, L.ptr(t), L.ptr(t), L.ptr(t), C.c_longlong(64), st)

    return graph_time(body) / n * 1e3


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    if os.environ.get("SMALL_N"):
        L.call("hg_set_option", b"small_n_tiles", int(os.environ["SMALL_N"]))
    if os.environ.get("SPLIT_AB"):
        # A/B in one process: weight tiles issued by a second thread (split_producer) vs. one producer thread
        def one(B, hw, cin, cout, k, flag):
            L.call("hg_set_option", b"split_producer", flag)
            torch.manual_seed(0)
            d = L.HgConvDesc(B, hw, hw, cin, cout, k, k, 1, k // 2, 1, L.HG_BF16)
            x = torch.randn(B, hw, hw, cin, device=dev).to(DT)
            y = torch.zeros(B, hw, hw, cout, device=dev, dtype=DT)
            w = (torch.randn(k * k, cout, cin, device=dev) * 0.05).to(DT)
            bias = torch.zeros(cout, device=dev)
            stats = torch.zeros(3 * cout, device=dev)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            L.call("hg_conv_fprop_ex", C.byref(d), L.ptr(x), L.ptr(w), L.ptr(bias), None, L.ptr(y), L.ptr(stats), None, st)
            torch.cuda.synchronize()
            return y, stats
        for hw in (4, 8, 16, 32, 64):
            for cin, cout, k in ((128, 128, 3), (256, 128, 1), (128, 256, 1)):
                y0, s0 = one(32, hw, cin, cout, k, 0)
                y1, s1 = one(32, hw, cin, cout, k, 1)
                same = torch.equal(y0, y1) and torch.allclose(s0, s1, rtol=1e-4)
                t = []
                for flag in (0, 1, 0, 1):
                    L.call("hg_set_option", b"split_producer", flag)
                    t.append(conv_chain(32, hw, cin, cout, k))
                print(f"@{hw}x{hw} conv{k}x{k} {cin}->{cout}: one producer {t[0]:.2f}/{t[2]:.2f} us, split {t[1]:.2f}/{t[3]:.2f} us, "
                      f"identical output: {same}", flush=True)
        sys.exit(0)
    if os.environ.get("WGRAD_BIG"):   # larger maps: input-channel split (fewer K splits -> fewer atomics, more dy re-reads)
        for hw in (32, 64):
            for cin, cout, k in ((128, 128, 3), (256, 128, 1), (128, 256, 1), (256, 256, 1)):
                t = []
                for npan in (0, 2, 1):
                    L.call("hg_set_option", b"wgrad_big_n_panels", npan)
                    t.append(wgrad_chain(32, hw, cin, cout, k))
                print(f"wgrad {k}x{k} {cin}->{cout} @{hw}x{hw}: all channels/CTA {t[0]:.2f} us | 128 ch/CTA {t[1]:.2f} | 64 ch/CTA {t[2]:.2f}",
                      flush=True)
        sys.exit(0)
    if os.environ.get("CONV_STATS"):   # cost of the per-channel statistics (smem column pass + global atomics) in fprop
        for hw in (4, 16, 32, 64):
            for cin, cout, k in ((128, 128, 3), (256, 128, 1), (128, 256, 1)):
                a, b = conv_chain(32, hw, cin, cout, k), conv_chain(32, hw, cin, cout, k, with_stats=False)
                print(f"@{hw}x{hw} conv{k}x{k} {cin}->{cout}: with statistics {a:.2f} us, without {b:.2f} us", flush=True)
        sys.exit(0)
    if os.environ.get("BN_GRID"):   # grid cap (blocks per SM) of the streaming BatchNorm kernels
        for hw in (8, 16):
            for per_sm in (1, 2, 6):
                L.call("hg_set_option", b"bn_bwd_blocks_per_sm", per_sm)
                print(f"@{hw} blocks/SM {per_sm}: bwd256+add {bn_bwd_chain(32, hw, 256, n=50, with_addend=True):.2f} "
                      f"bwd128 {bn_bwd_chain(32, hw, 128, n=50):.2f} us", flush=True)
        for per_sm in (1, 2, 6):
            L.call("hg_set_option", b"bn_blocks_per_sm", per_sm)
            L.call("hg_set_option", b"bn_bwd_blocks_per_sm", per_sm)
            r = []
            for hw in (32, 64):
                r += [bn_chain(32, hw, 256, n=20), bn_chain(32, hw, 128, n=20), bn_bwd_chain(32, hw, 256, n=20, with_addend=True),
                      bn_bwd_chain(32, hw, 128, n=20)]
            print(f"blocks/SM {per_sm:2d}: @32 apply256 {r[0]:.2f} apply128 {r[1]:.2f} bwd256+add {r[2]:.2f} bwd128 {r[3]:.2f} | "
                  f"@64 apply256 {r[4]:.2f} apply128 {r[5]:.2f} bwd256+add {r[6]:.2f} bwd128 {r[7]:.2f} us", flush=True)
        sys.exit(0)
    if os.environ.get("BN64"):   # the four BatchNorm streaming kernels at 64x64 only (ncu --set full capture)
        print(bn_chain(32, 64, 256, n=4), bn_chain(32, 64, 128, n=4), bn_bwd_chain(32, 64, 256, n=4, with_addend=True),
              bn_bwd_chain(32, 64, 128, n=4))
        sys.exit(0)
    if os.environ.get("WGRAD_T1"):
        def once(hw, cin, cout, k):
            torch.manual_seed(0)
            d = L.HgConvDesc(32, hw, hw, cin, cout, k, k, 1, k // 2, 1, L.HG_BF16)
            x = torch.randn(32, hw, hw, cin, device=dev).to(DT)
            dy = torch.randn(32, hw, hw, cout, device=dev).to(DT)
            dw = torch.zeros(k * k, cout, cin, device=dev)
            L.call("hg_conv_wgrad", C.byref(d), L.ptr(x), L.ptr(dy), L.ptr(dw), None,
                   C.c_void_p(torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            return dw
        for hw in (4, 8, 16, 32):
            for cin, cout, k in ((128, 128, 3), (256, 128, 1), (128, 256, 1), (256, 256, 1)):
                t, outs = [], []
                for thr, npan in ((0, 0), (128, 0), (128, 2), (128, 1)):
                    L.call("hg_set_option", b"wgrad_t1_max_kb", thr)
                    L.call("hg_set_option", b"wgrad_small_n_panels", npan)
                    outs.append(once(hw, cin, cout, k))
                    t.append(wgrad_chain(32, hw, cin, cout, k))
                err = max(((o - outs[0]).abs().max() / outs[0].abs().max()).item() for o in outs[1:])
                print(f"wgrad {k}x{k} {cin}->{cout} @{hw}x{hw}: base {t[0]:.2f} us | 1 tap/CTA {t[1]:.2f} | + 128-ch N tiles "
                      f"{t[2]:.2f} | + 64-ch N tiles {t[3]:.2f} | max rel diff {err:.1e}", flush=True)
        sys.exit(0)
    if os.environ.get("WGRAD"):
        # wgrad epilogue: per-thread red.v4 atomics vs shared-memory staging + cp.reduce.async.bulk (A/B in one process)
        def once(B, hw, cin, cout, k, bulk):
            L.call("hg_set_option", b"wgrad_bulk_reduce", bulk)
            torch.manual_seed(0)
            d = L.HgConvDesc(B, hw, hw, cin, cout, k, k, 1, k // 2, 1, L.HG_BF16)
            x = torch.randn(B, hw, hw, cin, device=dev).to(DT)
            dy = torch.randn(B, hw, hw, cout, device=dev).to(DT)
            dw = torch.zeros(k * k, cout, cin, device=dev)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            L.call("hg_conv_wgrad", C.byref(d), L.ptr(x), L.ptr(dy), L.ptr(dw), None, st)
            torch.cuda.synchronize()
            return dw
        for hw in (4, 8, 16, 32, 64):
            for cin, cout, k in ((128, 128, 3), (256, 128, 1), (128, 256, 1), (256, 256, 1), (256, 64, 1)):
                d0, d1 = once(32, hw, cin, cout, k, 0), once(32, hw, cin, cout, k, 1)
                err = ((d0 - d1).abs().max() / d0.abs().max()).item()
                t = []
                for bulk, dbg in ((0, 0), (1, 0), (1, 1), (1, 2)):
                    L.call("hg_set_option", b"wgrad_bulk_reduce", bulk)
                    L.call("hg_set_option", b"wgrad_dbg", dbg)
                    t.append(wgrad_chain(32, hw, cin, cout, k))
                L.call("hg_set_option", b"wgrad_dbg", 0)
                print(f"wgrad @{hw}x{hw} {k}x{k} {cin}->{cout}: atomics {t[0]:.2f} us | bulk reduce {t[1]:.2f} | staged, no reduce "
                      f"{t[2]:.2f} | no epilogue {t[3]:.2f} | rel diff {err:.1e}", flush=True)
        sys.exit(0)
    if os.environ.get("DBG_TS"):
        L.call("hg_set_option", b"dbg_ts", 1)
        for hw, cin, cout, k in ((4, 256, 128, 1), (4, 128, 128, 3), (64, 128, 128, 3)):
            t = conv_chain(32, hw, cin, cout, k, n=20)
            print(f"phases of CTA 0, conv {k}x{k} {cin}->{cout} @{hw}x{hw} (chain: {t:.2f} us/call)", flush=True)
            L.call("hg_set_option", b"dbg_ts", 2)
        L.call("hg_set_option", b"dbg_ts", 0)
        sys.exit(0)
    for opt in os.environ.get("HG_OPTIONS", "").split(","):
        if "=" in opt:
            k_, v_ = opt.split("=")
            L.call("hg_set_option", k_.encode(), int(v_))
    if os.environ.get("BN_ONLY"):
        for hw in (16, 32, 64):
            print(f"B=32 @{hw}x{hw}: bn_bwd_apply C256+add {bn_bwd_chain(32, hw, 256, with_addend=True):7.2f} us | "
                  f"bn_bwd_apply C128 {bn_bwd_chain(32, hw, 128):7.2f} us | bn_apply C256 {bn_chain(32, hw, 256):7.2f} us", flush=True)
        sys.exit(0)
    print(f"trivial kernel chain (hg_add 64 elems): {empty_chain():.2f} us/call")
    for hw in (4, 8, 16, 32, 64):
        print(f"B=32 @{hw}x{hw}: conv3x3 128->128 {conv_chain(32, hw, 128, 128, 3):7.2f} us | "
              f"conv1x1 256->128 {conv_chain(32, hw, 256, 128, 1):7.2f} us | "
              f"conv1x1 128->256 {conv_chain(32, hw, 128, 256, 1):7.2f} us | "
              f"conv1x1 256->256 {conv_chain(32, hw, 256, 256, 1):7.2f} us | "
              f"bn_apply C256 {bn_chain(32, hw, 256):7.2f} us | bn_apply C128 {bn_chain(32, hw, 128):7.2f} us | "
              f"bn_bwd_apply C256+add {bn_bwd_chain(32, hw, 256, with_addend=True):7.2f} us | "
              f"bn_bwd_apply C128 {bn_bwd_chain(32, hw, 128):7.2f} us")
