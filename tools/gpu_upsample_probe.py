"""Up-sampling kernels at the 32x32 -> 64x64 level (B=32, 256 channels): separable vs gather forms, grid caps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

B, DEV, BF = 32, "cuda", torch.bfloat16


def run(h, Cc, opts):
    for k, v in opts.items():
        L.call("hg_set_option", k.encode(), v)
    NR = 6
    low = [torch.randn(B, h, h, Cc, device=DEV).to(BF) for _ in range(NR)]
    skip = [torch.randn(B, 2 * h, 2 * h, Cc, device=DEV).to(BF) for _ in range(NR)]
    out = [torch.empty(B, 2 * h, 2 * h, Cc, device=DEV, dtype=BF) for _ in range(NR)]
    dlow = [torch.empty(B, h, h, Cc, device=DEV, dtype=BF) for _ in range(NR)]
    stats = torch.zeros(3 * Cc, device=DEV)
    st = L.stream_ptr()
    f = lambda i: L.call("hg_upsample2x_add_fwd", L.HG_BF16, 0, L.ptr(low[i]), L.ptr(skip[i]), B, h, h, Cc, L.ptr(out[i]),
                         L.ptr(stats), st)
    g = lambda i: L.call("hg_upsample2x_bwd", L.HG_BF16, 0, L.ptr(skip[i]), None, B, h, h, Cc, L.ptr(dlow[i]), st)
    res = []
    for fn in (f, g):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            fn(i % NR)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 50)
    print(f"h={h} C={Cc} {opts}: fwd {res[0]:.2f} us  bwd {res[1]:.2f} us", flush=True)


if __name__ == "__main__":
    for h in (32, 16):
        run(h, 256, {"upsample_sep": 0, "upsample_fwd_cap": 0})
        for cap in (0, 3, 6, 8):
            run(h, 256, {"upsample_sep": 1, "upsample_fwd_cap": cap})
