"""HBM-bound kernels of the hot path at the sizes of BASELINE configs[1] (8-stack training, B=32) and configs[4]
(17-joint inference, B=32 per GPU): BatchNorm apply / backward apply, max-pool, up-sample + add, the 8-term MSE, the
cross-entropy heads, argmax decode, the PCKh sweep and Gaussian target rendering.

Every kernel is launched REPS times on rotating buffers larger than the 126 MB L2 (so each launch streams from HBM) and
timed with CUDA events on the launching stream; algorithmic bytes per launch / that time = achieved GB/s against the
measured copy peak (MEASURED_PEAKS.json).  Run the same command under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv
to get the DRAM traffic of the same launches; tools/ncu_hbm_summary.py joins both into
profiles/r02_hbm_kernels_ncu.summary.txt.  Writes gpurun_out/hbm_kernels_events.json (manifest + event timings).
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_process_for_human_pose_estimation_b200 as hg  # noqa: E402
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402

B = int(os.environ.get("B", "32"))
REPS = int(os.environ.get("REPS", "10"))
DEV = "cuda"
BF = torch.bfloat16
manifest = []


def nrot_for(nbytes):
    """rotating copies so that the working set exceeds the L2 several times over"""
    return max(2, min(12, int(4 * 126e6 / max(nbytes, 1)) + 1))


def run(label, kernel_regex, nbytes, fn, nrot):
    for i in range(2):
        fn(i % nrot)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):     # back to back on the launching stream: launch latency overlaps the previous kernel
        fn((i + 2) % nrot)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / REPS
    manifest.append({"label": label, "kernel": kernel_regex, "algorithmic_bytes": int(nbytes), "launches": REPS + 2,
                     "event_us": round(us, 2), "event_gbs": round(nbytes / us / 1e3, 1)})
    print(f"{label:58s} {us:9.2f} us {nbytes / 1e6:9.1f} MB {nbytes / us / 1e3:8.0f} GB/s", flush=True)


def bn_cases(H, Cc):
    M, cp = B * H * H, L.pad64(Cc)
    eb = M * cp * 2
    nrot = nrot_for(4 * eb)
    xs = [torch.randn(M, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    ys = [torch.empty(M, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    gs = [torch.randn(M, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    ad = [torch.randn(M, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    gamma, beta = torch.ones(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    stats, red = torch.zeros(3 * cp, device=DEV), torch.zeros(2 * cp, device=DEV)
    dg, db = torch.zeros(Cc, device=DEV), torch.zeros(Cc, device=DEV)
    d = L.HgBnDesc(M, Cc, L.HG_BF16, 1e-5, 1, 0)
    st = L.stream_ptr()
    L.call("hg_bn_stats", C.byref(d), L.ptr(xs[0]), L.ptr(stats), st)
    L.call("hg_bn_bwd_reduce", C.byref(d), L.ptr(gs[0]), L.ptr(xs[0]), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None,
           L.ptr(red), st)
    run(f"bn_apply C{Cc} @{H}x{H}", "bn_apply_kernel", 2 * eb,
        lambda i: L.call("hg_bn_apply", C.byref(d), L.ptr(xs[i]), L.ptr(stats), L.ptr(gamma), L.ptr(beta), None, None,
                         L.ptr(ys[i]), st), nrot)
    run(f"bn_bwd_apply C{Cc} @{H}x{H}", "bn_bwd_apply_kernel", 3 * eb,
        lambda i: L.call("hg_bn_bwd_apply", C.byref(d), L.ptr(gs[i]), L.ptr(xs[i]), L.ptr(stats), L.ptr(gamma),
                         L.ptr(beta), None, None, L.ptr(red), None, L.ptr(ys[i]), L.ptr(dg), L.ptr(db), None, st), nrot)
    run(f"bn_bwd_apply +addend C{Cc} @{H}x{H}", "bn_bwd_apply_kernel", 4 * eb,
        lambda i: L.call("hg_bn_bwd_apply", C.byref(d), L.ptr(gs[i]), L.ptr(xs[i]), L.ptr(stats), L.ptr(gamma),
                         L.ptr(beta), None, None, L.ptr(red), L.ptr(ad[i]), L.ptr(ys[i]), L.ptr(dg), L.ptr(db), None, st),
        nrot)


def spatial_cases(H, Cc):
    cp = L.pad64(Cc)
    big, small = B * H * H * cp * 2, B * (H // 2) * (H // 2) * cp * 2
    nrot = nrot_for(2 * big)
    xs = [torch.randn(B, H, H, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    ps = [torch.empty(B, H // 2, H // 2, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    gp = [torch.randn(B, H // 2, H // 2, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    os_ = [torch.empty(B, H, H, cp, device=DEV, dtype=BF) for _ in range(nrot)]
    stats = torch.zeros(3 * cp, device=DEV)
    st = L.stream_ptr()
    run(f"maxpool2_fwd (+stats) C{Cc} @{H}x{H}", "maxpool2_fwd", big + small,
        lambda i: L.call("hg_maxpool2_fwd", L.HG_BF16, L.ptr(xs[i]), B, H, H, Cc, L.ptr(ps[i]), L.ptr(stats), st), nrot)
    run(f"maxpool2_bwd C{Cc} @{H}x{H}", "maxpool2_bwd", 2 * big + small,
        lambda i: L.call("hg_maxpool2_bwd", L.HG_BF16, L.ptr(xs[i]), L.ptr(gp[i]), None, B, H, H, Cc, L.ptr(os_[i]), st), nrot)
    run(f"upsample2x_add_fwd bilinear (+stats) C{Cc} @{H // 2}->{H}", "upsample2_add_fwd|upsample2x", 2 * big + small,
        lambda i: L.call("hg_upsample2x_add_fwd", L.HG_BF16, 0, L.ptr(gp[i]), L.ptr(xs[i]), B, H // 2, H // 2, Cc,
                         L.ptr(os_[i]), L.ptr(stats), st), nrot)
    run(f"upsample2x_bwd bilinear C{Cc} @{H}->{H // 2}", "upsample2_bwd|upsample2x", big + small,
        lambda i: L.call("hg_upsample2x_bwd", L.HG_BF16, 0, L.ptr(xs[i]), None, B, H // 2, H // 2, Cc, L.ptr(ps[i]), st),
        nrot)


def loss_cases():
    J, S = 16, 8
    n = B * J * 64 * 64
    nrot = 3
    outs = [[torch.randn(B, J, 64, 64, device=DEV, requires_grad=True) for _ in range(S)] for _ in range(nrot)]
    tg = [torch.rand(B, J, 64, 64, device=DEV) for _ in range(nrot)]

    def mse(i):
        hg.mse_losses(outs[i], tg[i]).sum().backward()   # forward (losses + gradients in one sweep) and the hand-off
    run(f"mse_multi 8 stacks [{B},16,64,64] fwd+grad", "mse_multi_kernel", (2 * S + 1) * n * 4, mse, nrot)
    Cc = 18
    lg = [torch.randn(B, Cc, 64, 64, device=DEV, requires_grad=True) for _ in range(nrot)]
    lb = [torch.randint(0, Cc, (B, 64, 64), device=DEV) for _ in range(nrot)]

    def ce(i):
        hg.cross_entropy_losses([(lg[i], lb[i])]).sum().backward()
    run(f"ce_multi 1 term [{B},18,64,64] fwd+grad", "ce_multi_kernel|ce_count_kernel", 2 * B * Cc * 4096 * 4 + B * 4096 * 8,
        ce, nrot)


def eval_cases():
    J = 17
    nrot = 4
    hm = [torch.randn(B, J, 64, 64, device=DEV) for _ in range(nrot)]
    hmh = [h.half() for h in hm]
    run(f"decode_argmax fp32 [{B},17,64,64]", "decode_argmax", B * J * 4096 * 4, lambda i: hg.decode_argmax(hm[i]), nrot)
    run(f"decode_argmax fp16 [{B},17,64,64]", "decode_argmax", B * J * 4096 * 2, lambda i: hg.decode_argmax(hmh[i]), nrot)
    r = np.random.RandomState(7)
    label = torch.zeros(B, 64, 64, dtype=torch.int64)
    for b_ in range(B):
        pos = r.choice(64 * 64, J, replace=False)
        for j, pp in enumerate(pos):
            label[b_, pp // 64, pp % 64] = j + 1
    label = label.to(DEV)
    x0 = r.uniform(5, 40, [B, 2]).astype("float32")
    rect = torch.from_numpy(np.concatenate([x0, x0 + r.uniform(5, 20, [B, 2]).astype("float32")], 1)).to(DEV)
    run(f"pckh_sweep fp32 [{B},17,64,64] + labels", "pckh_sweep", B * J * 4096 * 4 + B * 4096 * 8,
        lambda i: hg.pckh_sweep_counts(hm[i], label, rect, 0), nrot)
    run(f"softmax_stats + pckh_sweep on logits [{B},17,64,64]", "softmax_stats|pckh_sweep", 2 * B * J * 4096 * 4 + B * 4096 * 16,
        lambda i: hg.pckh_sweep_counts(hm[i], label, rect, 1, njoints=16, logits=True), nrot)
    kp = np.zeros([B, 1, 16, 3])
    kp[..., 0], kp[..., 1], kp[..., 2] = r.randint(0, 640, [B, 1, 16]), r.randint(0, 480, [B, 1, 16]), 2
    wh = np.tile(np.array([[640.0, 480.0]]), (B, 1))
    run(f"render_gauss [{B},16,64,64] (write)", "render_gauss", B * 16 * 4096 * 4,
        lambda i: hg.gaussian_heatmaps(kp, wh, truncate=True, device=DEV), 1)


def main():
    print(torch.cuda.get_device_name(0), f"B={B} REPS={REPS}")
    for H, Cc in ((64, 256), (64, 128), (32, 256), (32, 128)):
        bn_cases(H, Cc)
    for H in (64, 32):
        spatial_cases(H, 256)
    loss_cases()
    eval_cases()
    os.makedirs("gpurun_out", exist_ok=True)
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
    json.dump({"B": B, "reps": REPS, "hbm_gbs_peak": peaks["hbm_gbs"], "kernels": manifest},
              open("gpurun_out/hbm_kernels_events.json", "w"), indent=1)


if __name__ == "__main__":
    main()
