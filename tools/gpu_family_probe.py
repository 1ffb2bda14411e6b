"""Bring-up probe: every mirrored model family against its reference golden, one subprocess per family."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(script, factory, dtype, mode="train"):
    import numpy as np
    import torch

    import progressive_process_for_human_pose_estimation_b200 as hg

    g = np.load(os.path.join(ROOT, "tests", "golden", f"family_{script}_{mode}.npz"))
    hg.set_compute_dtype(getattr(torch, dtype))
    mod = importlib.import_module(f"progressive_process_for_human_pose_estimation_b200.{script}")
    torch.manual_seed(0)
    net = getattr(mod, factory)()
    if mode == "eval":
        from tests.test_families import randomize_running_stats
        randomize_running_stats(net)
        net.eval()
    net = net.cuda()
    gen = torch.Generator().manual_seed(int(g["input_seed"]))
    x = torch.randn(int(g["B"]), 3, int(g["S"]), int(g["S"]), generator=gen)
    out = net(x.cuda())
    tgts = [torch.rand(o.shape, generator=gen) for o in out]
    losses = [torch.nn.functional.mse_loss(o, t.cuda()) for o, t in zip(out, tgts)]
    sum(losses).backward()
    torch.cuda.synchronize()
    for i, o in enumerate(out):
        ref = g[f"out{i}"]
        a = o.detach().cpu().numpy().astype(np.float64)
        print(f"  out{i} rel {np.linalg.norm(a - ref) / np.linalg.norm(ref):.3e}  loss {losses[i].item():.6f} vs {g['losses'][i]:.6f}")
    names = [str(n) for n in g["param_names"]]
    params = dict(net.named_parameters())
    worst = []
    for i, n in enumerate(names):
        p = params[n]
        if g["grad_is_none"][i]:
            if p.grad is not None and p.grad.abs().max().item() != 0:
                print("  UNEXPECTED grad", n)
            continue
        if p.grad is None:
            print("  MISSING grad", n)
            continue
        if g["grad_norm"][i] < 1e-4 * np.median(g["grad_norm"][g["grad_norm"] > 0]):
            continue
        f = p.grad.detach().double().reshape(-1).cpu()
        idx = torch.linspace(0, f.numel() - 1, 24).long()
        d = f[idx].numpy()
        r = g["grad_digest"][i][2:]
        worst.append((np.linalg.norm(d - r) / (np.linalg.norm(r) + 1e-30), n))
    worst.sort(reverse=True)
    print("  worst grads:", [(f"{w[0]:.2e}", w[1]) for w in worst[:6]], "median", f"{worst[len(worst) // 2][0]:.2e}")
    sd = net.state_dict()
    keys = [str(k) for k in g["keys"]]
    bad = 0
    for i, k in enumerate(keys):
        if "num_batches_tracked" in k and float(sd[k]) != g["after_digest"][i][2]:
            bad += 1
    print("  num_batches_tracked mismatches:", bad)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        FAMILIES = [("try_different_stack", "creatModel"), ("try_different_stack_without_skeleton", "creatModel"),
                    ("try_with_aspp", "creatModel"), ("try_with_aspp_remove_max_pool", "creatModel"),
                    ("try_skeleton_and_keypoints", "creatModel"), ("hourglass_compare", "creatModel")]

        for script, factory in FAMILIES:
            for dtype, mode in (("float32", "eval"), ("bfloat16", "eval")):
                print(f"== {script} {dtype} {mode}", flush=True)
                env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
                r = subprocess.run([sys.executable, __file__, script, factory, dtype, mode], env=env, capture_output=True, text=True)
                print(r.stdout[-1500:], r.stderr[-1200:] if r.returncode else "", flush=True)
