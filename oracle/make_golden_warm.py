"""TEST INFRASTRUCTURE ONLY -- warm-weight golden of the HEADLINE configuration (BASELINE.json configs[1]:
try_with_torch.creatModel, nStack=8, 16 heatmaps), SURVEY Q13 protocol (ii).

At seeded random init the weight-shared 8-stack network is numerically chaotic in train mode (fp32 vs fp64 differ by
12 % at stack 8), so a train-mode end-to-end comparison is only meaningful on WARM weights.  This script executes the
REAL reference (/root/reference/try_with_torch.py through oracle/refload.py):

  1. seeded init, K = 80 fp32 Adam steps (lr 1e-3, B=4, Gaussian targets) of the reference's own training step
     (try_with_torch.py:329-344 with eight loss terms);
  2. the resulting state_dict is rounded to bf16-representable values (so that the fixture is half the size and both
     compute paths start from identical numbers) and loaded back into the reference;
  3. ONE train-mode step of the reference on a fresh batch (B=2): per-stack heatmaps, per-stack MSE, every parameter
     gradient, the BatchNorm buffers after the step; and the same forward under CPU autocast(bfloat16) as the yardstick
     for what bf16 arithmetic does to this network.

    python -m oracle.make_golden_warm      # rewrites tests/golden/warm_s_8stack.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload, targets_np  # noqa: E402
from oracle.make_golden import GOLDEN, digest  # noqa: E402

K_STEPS = int(os.environ.get("HG_WARM_STEPS", "80"))
NSTACK, J = 8, 16


def gauss_targets(seed, B):
    r = np.random.RandomState(seed)
    kp = np.zeros([B, 1, J, 3])
    kp[..., 0] = r.randint(0, 640, [B, 1, J])
    kp[..., 1] = r.randint(0, 480, [B, 1, J])
    kp[..., 2] = r.randint(0, 3, [B, 1, J])
    maps = np.stack([targets_np.gauss_map(kp[b], (640.0, 480.0), J, truncate=True) for b in range(B)])
    return torch.from_numpy(maps), kp


def bf16_round(t):
    return t.bfloat16().float()


def bf16_bits(t):
    return t.bfloat16().view(torch.int16).numpy().astype(np.uint16)


def main():
    ref = refload.load("try_with_torch")
    ref.nStack, ref.nOutChannels = NSTACK, J
    torch.manual_seed(0)
    net = ref.creatModel()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    mse = [torch.nn.MSELoss() for _ in range(NSTACK)]
    g = torch.Generator().manual_seed(11)
    warm_losses = []
    for it in range(K_STEPS):
        x = torch.randn(4, 3, 256, 256, generator=g)
        y, _ = gauss_targets(100 + it, 4)
        out = net(x)
        loss = sum(mse[k].forward(out[k], y) for k in range(NSTACK))
        opt.zero_grad()
        loss.backward()
        opt.step()
        warm_losses.append(loss.item())
        print(f"warm step {it}: loss {loss.item():.5f}", flush=True)
    # ---- warm state, bf16-representable ---------------------------------------------------------------
    sd = {k: (bf16_round(v) if v.is_floating_point() else v.clone()) for k, v in net.state_dict().items()}
    net.load_state_dict(sd)
    keys = list(sd.keys())
    x = bf16_round(torch.randn(2, 3, 256, 256, generator=g))
    y, kp = gauss_targets(999, 2)
    # yardstick first (no_grad, does not touch the weights; BN buffers restored afterwards)
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        oa = [o.float() for o in net(x)]
    # fp64 yardstick: the reference's own fp32-vs-fp64 divergence on these weights
    net.double()
    net.zero_grad(set_to_none=True)
    o64 = net(x.double())
    sum(torch.nn.functional.mse_loss(o, y.double()) for o in o64).backward()
    g64 = {n: (p.grad.clone() if p.grad is not None else None) for n, p in net.named_parameters()}
    o64 = [o.detach() for o in o64]
    net.float()
    net.load_state_dict(sd)
    net.zero_grad(set_to_none=True)
    out = net(x)
    per = [mse[k].forward(out[k], y) for k in range(NSTACK)]
    sum(per).backward()
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()  # noqa: E731
    yard = [rel(oa[k], out[k].detach()) for k in range(NSTACK)]
    yard64 = [rel(out[k].detach(), o64[k]) for k in range(NSTACK)]
    print("fp32 vs fp64 rel-L2 per stack", ["%.2e" % v for v in yard64])
    print("per-stack loss", [round(p.item(), 6) for p in per])
    print("autocast-bf16 vs fp32 rel-L2 per stack", [round(v, 4) for v in yard])
    pnames = [n for n, _ in net.named_parameters()]
    small = [n for n, p in net.named_parameters() if p.grad is not None and p.numel() <= 4096]
    after = net.state_dict()
    np.savez_compressed(
        os.path.join(GOLDEN, "warm_s_8stack.npz"),
        k_steps=K_STEPS, nStack=NSTACK, J=J, warm_losses=np.array(warm_losses),
        keys=np.array(keys),
        **{"w:" + k: (bf16_bits(v) if v.is_floating_point() else v.numpy()) for k, v in sd.items()},
        x_bits=bf16_bits(x), target=y.numpy(), keypoints=kp,
        out_digest=np.stack([digest(o) for o in out]),
        out0=out[0].detach().numpy(), out3=out[3].detach().numpy(), out7=out[7].detach().numpy(),
        losses=np.array([p.item() for p in per], dtype=np.float64),
        yardstick_bf16=np.array(yard), yardstick_fp64=np.array(yard64),
        param_names=np.array(pnames),
        grad_is_none=np.array([p.grad is None for _, p in net.named_parameters()]),
        grad_digest=np.stack([digest(p.grad) if p.grad is not None else np.zeros(26) for _, p in net.named_parameters()]),
        grad_norm=np.array([p.grad.double().norm().item() if p.grad is not None else 0.0
                            for _, p in net.named_parameters()]),
        # the reference's own fp32-vs-fp64 divergence per gradient tensor (yardstick for the fp32 path)
        grad_noise_fp64=np.array([rel(p.grad, g64[n]) if p.grad is not None and g64[n] is not None and
                                  g64[n].norm() > 0 else 0.0 for n, p in net.named_parameters()]),
        small_grad_names=np.array(small),
        **{"g:" + n: dict(net.named_parameters())[n].grad.numpy() for n in small},
        after_digest=np.stack([digest(after[k].float()) for k in keys]),
    )
    print("wrote", os.path.join(GOLDEN, "warm_s_8stack.npz"),
          os.path.getsize(os.path.join(GOLDEN, "warm_s_8stack.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
