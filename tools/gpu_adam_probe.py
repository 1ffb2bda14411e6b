"""GPU and host time of one optimizer step over the 199 parameter tensors of the 8-stack model: hg.Adam (one launch)
vs torch.optim.Adam (foreach) vs torch.optim.Adam(fused=True)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_process_for_human_pose_estimation_b200 as hg  # noqa: E402
import progressive_process_for_human_pose_estimation_b200.try_with_torch as m  # noqa: E402

m.nStack, m.nOutChannels = 8, 16
net = m.creatModel().cuda()
for p in net.parameters():
    p.grad = torch.randn_like(p)
for name, mk in (("hg.Adam", lambda ps: hg.Adam(ps, lr=1e-4)), ("torch foreach", lambda ps: torch.optim.Adam(ps, lr=1e-4)),
                 ("torch fused", lambda ps: torch.optim.Adam(ps, lr=1e-4, fused=True))):
    opt = mk(list(net.parameters()))
    for _ in range(3):
        opt.step()
    torch.cuda.synchronize()
    torch.cuda._sleep(int(2e8))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        opt.step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:14s}: host {1e3 * (t1 - t0) / 20:.3f} ms/step, device span {e0.elapsed_time(e1) / 20:.3f} ms/step")
