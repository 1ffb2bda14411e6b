"""Bring-up probe for the whole-network plan (run on a B200 through gpurun): product path vs the CPU oracle."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_process_for_human_pose_estimation_b200 as hg  # noqa: E402
import progressive_process_for_human_pose_estimation_b200.try_with_torch as m  # noqa: E402
from oracle import hourglass_torch as ho  # noqa: E402


def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def run(dtype, nStack, B, S, steps=3):
    m.nStack = nStack
    m.nOutChannels = 16
    hg.set_compute_dtype(dtype)
    torch.manual_seed(0)
    net = m.creatModel()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 3, S, S, generator=g)
    tgt = torch.rand(B, 16, S // 4, S // 4, generator=g)
    # oracle (CPU fp32)
    sd = ho.clone_state(sd0, requires_grad=True)
    cfg = ho.Config(nStack=nStack, nOutChannels=16)
    t0 = time.time()
    oo = ho.creat_model_s(sd, x, cfg)
    tot, _ = ho.mse_losses(oo, tgt)
    tot.backward()
    t_or = time.time() - t0
    xc, tc = x.cuda(), tgt.cuda()
    for it in range(steps):
        for p in net.parameters():
            p.grad = None
        if it > 0:
            net.load_state_dict(sd0)  # restore BN running stats so every iteration is the same computation
        out = net(xc)
        loss = sum(torch.nn.MSELoss()(o, tc) for o in out)
        loss.backward()
        torch.cuda.synchronize()
        print(f"[{dtype} nStack={nStack} B={B} S={S}] iter {it}: loss {loss.item():.6f} (oracle {tot.item():.6f})")
        for k, (a, b) in enumerate(zip(out, oo)):
            print(f"   out[{k}] rel-L2 {rel(a.cpu(), b.detach()):.3e}  max|ref| {b.abs().max().item():.3e}")
        worst = []
        for name, p in net.named_parameters():
            go = sd[name].grad
            if p.grad is None:
                assert go is None or go.abs().max() == 0, name
                continue
            worst.append((rel(p.grad.cpu(), go), name, go.norm().item()))
        worst.sort(reverse=True)
        print("   worst grads:", [(f"{w[0]:.2e}", w[1], f"{w[2]:.1e}") for w in worst[:6]])
        big = [w for w in worst if w[2] > 1e-6]
        print("   median grad rel err:", sorted(w[0] for w in big)[len(big) // 2])
        sdn = net.state_dict()
        rs = max(rel(sdn[k].float().cpu(), sd[k].detach().float()) for k in sdn if "running" in k)
        nb = all(int(sdn[k]) == int(sd[k]) for k in sdn if "num_batches" in k)
        print(f"   running stats worst rel {rs:.3e}, num_batches_tracked equal {nb}")
    print("   launches fwd/bwd:", net.launches_per_step(), "oracle cpu time %.2fs" % t_or)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run(torch.float32, 1, 2, 256)
    run(torch.float32, 2, 3, 256)
    run(torch.bfloat16, 1, 2, 256)
    run(torch.bfloat16, 2, 4, 256)
