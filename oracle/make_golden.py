"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the REAL reference code from
/root/reference (through oracle/refload.py) on seeded synthetic inputs.  The fixtures travel to the GPU box,
where /root/reference does not exist; the oracle restatements and the CUDA path are checked against them.

    python -m oracle.make_golden          # rewrites tests/golden/
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refload  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
SAMPLE = 24  # sampled entries per tensor in digests


def digest(t):
    """(sum, abs-sum, first SAMPLE strided entries) of a tensor, in float64."""
    f = t.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, SAMPLE).long()
    return np.concatenate([[f.sum().item(), f.abs().sum().item()], f[idx].numpy()])


def model_inputs(seed, B, S, J):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, S, S, generator=g)
    tgt = torch.rand(B, J, S // 4, S // 4, generator=g)
    return x, tgt


def golden_model_s():
    """try_with_torch.creatModel, nStack=2, 16 heatmaps, B=2, 128x128 input: forward, 2x MSE, backward."""
    ref = refload.load("try_with_torch")
    ref.nStack, ref.nOutChannels = 2, 16
    torch.manual_seed(0)
    net = ref.creatModel()
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    x, tgt = model_inputs(1, 2, 128, 16)
    out = net(x)
    l1 = torch.nn.MSELoss().forward(out[0], tgt)
    l2 = torch.nn.MSELoss().forward(out[1], tgt)
    (l1 + l2).backward()
    keys = list(sd0.keys())
    pnames = [n for n, _ in net.named_parameters()]
    np.savez_compressed(
        os.path.join(GOLDEN, "model_s_2stack.npz"),
        seed=0, input_seed=1, B=2, S=128, J=16, nStack=2,
        keys=np.array(keys), param_names=np.array(pnames),
        state_digest=np.stack([digest(sd0[k].float()) for k in keys]),
        out0=out[0].detach().numpy(), out1=out[1].detach().numpy(),
        losses=np.array([l1.item(), l2.item()], dtype=np.float64),
        grad_is_none=np.array([p.grad is None for _, p in net.named_parameters()]),
        grad_digest=np.stack([digest(p.grad) if p.grad is not None else np.zeros(SAMPLE + 2)
                              for _, p in net.named_parameters()]),
        after_digest=np.stack([digest(net.state_dict()[k].float()) for k in keys]),
    )
    ref.nStack, ref.nOutChannels = 4, 17


def golden_model_c1():
    """BASELINE config 1: only_one_hourgless.creatModel (1 stack), 16 heatmaps, B=2, 256x256, forward + MSE."""
    ref = refload.load("only_one_hourgless")
    old = ref.nOutChannels
    ref.nOutChannels = 16
    torch.manual_seed(0)
    net = ref.creatModel()
    x, tgt = model_inputs(2, 2, 256, 16)
    out = net(x)
    loss = torch.nn.MSELoss()(out[0], tgt)
    np.savez_compressed(os.path.join(GOLDEN, "model_c1_1stack.npz"), seed=0, input_seed=2, B=2, S=256, J=16,
                        out0=out[0].detach().numpy().astype(np.float32), loss=np.float64(loss.item()))
    ref.nOutChannels = old


class FakeCOCO:
    """Minimal pycocotools.coco.COCO stand-in serving synthetic annotations to the reference datasets."""

    persons = {}
    skeleton = None

    def __init__(self, anno):
        pass

    def getCatIds(self):
        return [1]

    def getImgIds(self, catIds=None):
        return sorted(self.persons.keys())

    def loadImgs(self, i):
        return [{"file_name": "img.jpg"}]

    def getAnnIds(self, i):
        return [(i, p) for p in range(len(self.persons[i]))]

    def loadAnns(self, ids):
        return [{"keypoints": self.persons[i][p], "category_id": 1} for (i, p) in ids]

    def loadCats(self, c):
        return [{"skeleton": self.skeleton}]


def golden_targets():
    """myImageDataset_COCO.__getitem__ of try_different_stack.py (Gaussians + skeleton + background maps) and of
    try_skeleton_and_keypoints.py (keypoint + skeleton maps) on synthetic annotations (640x480 image)."""
    from PIL import Image

    tds = refload.load("try_different_stack")
    tsk = refload.load("try_skeleton_and_keypoints")
    r = np.random.RandomState(0)
    W, H = 640, 480
    n_img, P, J = 6, 3, 17
    kp = np.zeros([n_img, P, J, 3])
    kp[..., 0] = r.randint(0, W, [n_img, P, J])
    kp[..., 1] = r.randint(0, H, [n_img, P, J])
    kp[..., 2] = r.randint(0, 3, [n_img, P, J])
    npers = r.randint(1, P + 1, n_img)
    sks1 = (np.array(tds.sks) + 1).tolist() if hasattr(tds, "sks") else None
    if sks1 is None:
        tw = refload.load("try_with_torch")
        sks1 = (np.array(tw.sks) + 1).tolist()
    FakeCOCO.skeleton = sks1
    FakeCOCO.persons = {i: [kp[i, p].reshape(-1).astype(np.int64).tolist() for p in range(npers[i])]
                        for i in range(n_img)}
    tmp = tempfile.mkdtemp()
    Image.fromarray(np.zeros([H, W, 3], dtype=np.uint8)).save(os.path.join(tmp, "img.jpg"))
    tr = lambda im: torch.zeros(1)  # noqa: E731  (the image tensor is not part of the fixture)
    tds.COCO = FakeCOCO
    tsk.COCO = FakeCOCO
    d1 = tds.myImageDataset_COCO("x", tmp, tr)
    d2 = tsk.myImageDataset_COCO("x", tmp, tr)
    gauss, skel, bg, kpm, skel2 = [], [], [], [], []
    for i in range(n_img):
        _, g, s, b = d1[i]
        gauss.append(g.numpy())
        skel.append(s.numpy())
        bg.append(b.numpy())
        _, k, s2 = d2[i]
        kpm.append(k.numpy())
        skel2.append(s2.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "targets_coco.npz"), keypoints=kp, num_persons=npers.astype(np.int32),
                        img_wh=np.tile(np.array([[W, H]], dtype=np.float64), (n_img, 1)),
                        limbs=np.array(sks1) - 1, gauss=np.stack(gauss), skeleton=np.stack(skel),
                        background=np.stack(bg), keypoint_map=np.stack(kpm), skeleton2=np.stack(skel2))


def golden_pckh():
    """PCKh A/B/C of the reference on random + adversarial heatmaps."""
    hc = refload.load("hourglass_compare")
    pc = refload.load("performance_compare")
    oo = refload.load("only_one_hourgless")
    pc.nKeypoint_MPII = 16
    from oracle.synth import pckh_inputs

    d = pckh_inputs(0)
    B = d["x"].shape[0]
    tgt, rect = torch.from_numpy(d["target"]), torch.from_numpy(d["rect"])
    acc_c, pred_c, lab_c = hc.PCKh().forward(torch.from_numpy(d["x"]), tgt, rect)
    acc_b, pred_b, lab_b, std_b = pc.PCKh().forward(torch.from_numpy(d["x17"]), tgt, rect)
    oo.batch_size = B
    acc_a = oo.PCKh().forward(torch.from_numpy(d["x14"]), torch.from_numpy(d["t14"]))
    np.savez_compressed(os.path.join(GOLDEN, "pckh.npz"), seed=0, acc_c=acc_c, pred_c=np.stack(pred_c),
                        lab_c=np.stack(lab_c), acc_b=acc_b, pred_b=np.stack(pred_b), lab_b=np.stack(lab_b),
                        std_b=np.array([float(s) for s in std_b], dtype=np.float32), acc_a=np.float64(acc_a))


def golden_pckh_d():
    """PCKh D (calculate_parameters.py:906-937) and, on the same near-label inputs, PCKh B of the reference."""
    cp = refload.load("calculate_parameters")
    pc = refload.load("performance_compare")
    pc.nKeypoint_MPII = 16
    from oracle.synth import pckh_near_inputs

    d = pckh_near_inputs(0)
    x, tgt, rect = torch.from_numpy(d["x17"]), torch.from_numpy(d["target"]), torch.from_numpy(d["rect"])
    acc_d, pred_d, lab_d = cp.PCKh().forward(x, tgt, rect)
    acc_b, pred_b, lab_b, _ = pc.PCKh().forward(x, tgt, rect)
    np.savez_compressed(os.path.join(GOLDEN, "pckh_d.npz"), seed=0, acc_d=np.array(acc_d, dtype=np.float64),
                        pred_d=np.stack(pred_d), lab_d=np.stack(lab_d), acc_b=acc_b, pred_b=np.stack(pred_b))


def golden_aspp():
    """train.ASPP_Block / train._ASPPModule of the reference (train.py:449-495): seeded weights, train-mode forward +
    backward and eval-mode forward on [4,256,8,8]."""
    tr = refload.load("train")
    torch.manual_seed(0)
    net = tr.ASPP_Block()
    with torch.no_grad():   # non-trivial BatchNorm parameters / running statistics
        g = torch.Generator().manual_seed(5)
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.2)
                m.running_mean.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.bias.shape, generator=g) + 0.5)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(6)
    x = torch.randn(4, 256, 8, 8, generator=g).requires_grad_()
    w = torch.randn(4, 256, 8, 8, generator=g)
    net.train()
    out = net(x)
    (out * w).sum().backward()
    # yardstick: how far the reference's OWN bf16-autocast run is from its fp32 run on this input (BatchNorm over the
    # 4 samples of the image-level branch amplifies rounding in the backward pass)
    net_ac = tr.ASPP_Block()
    net_ac.load_state_dict(sd0)
    net_ac.train()
    x_ac = x.detach().clone().requires_grad_()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_ac = net_ac(x_ac)
    (out_ac.float() * w).sum().backward()
    noise = np.array([((out_ac.float() - out).norm() / out.norm()).item(), ((x_ac.grad - x.grad).norm() / x.grad.norm()).item()])
    named = list(net.named_parameters())
    # per-parameter gradient digests (norm and a seeded random projection) keep the fixture small
    gp = torch.Generator().manual_seed(7)
    arrays = {"grad_names": np.array([k for k, _ in named]),
              "grad_norm": np.array([p.grad.double().norm().item() for _, p in named]),
              "grad_proj": np.array([(p.grad.double().flatten() * torch.randn(p.numel(), generator=gp).double()).sum().item()
                                     for _, p in named])}
    sd1 = {k: v.clone() for k, v in net.state_dict().items()}
    net.load_state_dict(sd0)
    net.eval()
    with torch.no_grad():
        out_eval = net(x.detach())
    torch.manual_seed(1)
    am = tr._ASPPModule(256, 256, 3, padding=6, dilation=6)
    am.train()
    out_am = am(x.detach())
    f16 = lambda t: t.detach().numpy().astype(np.float16)  # noqa: E731  (reference values to 1e-3: tolerance is 1e-2)
    np.savez_compressed(os.path.join(GOLDEN, "aspp_block.npz"), out_train=out.detach().numpy(), gx=x.grad.numpy(),
                        out_eval=f16(out_eval), out_module=f16(out_am), noise_bf16_autocast=noise,
                        running_mean_after=sd1["conv1.1.running_mean"].numpy(),
                        gap_running_var_after=sd1["global_avg_pool.2.running_var"].numpy(), **arrays)


# (reference script, model factory attribute, drop-in module name)
FAMILIES = [
    ("try_different_stack", "creatModel"),
    ("try_different_stack_without_skeleton", "creatModel"),
    ("try_with_aspp", "creatModel"),
    ("try_with_aspp_remove_max_pool", "creatModel"),
    ("try_skeleton_and_keypoints", "creatModel"),
    ("hourglass_compare", "creatModel"),  # = performance_compare.creatModel_hourglass (same network, same fixture)
    ("try_more_layer", "creatModel"),     # executed inline ASPP at the bottom level, 4 stacks, last head reused
    ("train", "creatModel"),              # progressive model: Q4 blocks, stride-2 down-sampling, ASPP bottom, cat skips
    ("try_skeleton_from_keypoints_merge", "creatModel"),   # gather-add limb maps: 17-ch head -> 36-ch output (N3)
]


def family_inputs(seed, B, S):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 3, S, S, generator=g), g


def randomize_running_stats(net, seed=5):
    """Non-trivial BatchNorm running statistics for the eval-mode goldens (seeded; same recipe in the tests)."""
    g = torch.Generator().manual_seed(seed)
    for mod in net.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)


def golden_families(only=None):
    """One seeded fp32 step (forward, sum of per-output MSE against uniform random targets, backward) of every
    other model family the drop-in mirrors, executed by the REFERENCE classes: outputs, losses, digests of every
    gradient and of the state_dict before / after (BN running statistics).  B=2, 128x128 input.  Two variants:
    `train` (batch statistics: chaotic at random init, SURVEY Q13 -- checked against the fp64 yardstick stored with
    it) and `eval` (seeded running statistics: well conditioned, so forward AND backward parity are tight)."""
    for script, factory in FAMILIES:
        if only is not None and script not in only:
            continue
        for mode in ("train", "eval"):
            ref = refload.load(script)
            torch.manual_seed(0)
            net = getattr(ref, factory)()
            if mode == "eval":
                randomize_running_stats(net)
                net.eval()
            sd0 = {k: v.clone() for k, v in net.state_dict().items()}
            x, g = family_inputs(11, 2, 128)
            out = net(x)
            tgts = [torch.rand(o.shape, generator=g) for o in out]
            losses = [torch.nn.functional.mse_loss(o, t) for o, t in zip(out, tgts)]
            sum(losses).backward()
            keys = list(sd0.keys())
            named = list(net.named_parameters())
            arrays = {f"out{i}": o.detach().numpy() for i, o in enumerate(out)}
            # yardstick: the reference's OWN fp32-vs-fp64 divergence on this step -- what the rounding
            # noise of a correct implementation amounts to after the network amplified it
            torch.manual_seed(0)
            net64 = getattr(ref, factory)()
            net64.load_state_dict(sd0)
            net64 = net64.double()
            if mode == "eval":
                net64.eval()
            out64 = net64(x.double())
            sum(torch.nn.functional.mse_loss(o, t.double()) for o, t in zip(out64, tgts)).backward()
            relf = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))  # noqa: E731
            out_noise = np.array([relf(a.detach(), b.detach()) for a, b in zip(out, out64)])
            grad_noise = np.array([relf(p.grad, q.grad) if p.grad is not None else 0.0
                                   for (_, p), (_, q) in zip(named, net64.named_parameters())])
            # second yardstick: the reference's own divergence when IT computes in bf16 (torch.autocast on CPU)
            torch.manual_seed(0)
            net16 = getattr(ref, factory)()
            net16.load_state_dict(sd0)
            if mode == "eval":
                net16.eval()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                out16 = net16(x)
            sum(torch.nn.functional.mse_loss(o.float(), t) for o, t in zip(out16, tgts)).backward()
            out_noise16 = np.array([relf(a.detach().float(), b.detach()) for a, b in zip(out16, out64)])
            grad_noise16 = np.array([relf(p.grad, q.grad) if p.grad is not None else 0.0
                                     for (_, p), (_, q) in zip(net16.named_parameters(), net64.named_parameters())])
            arrays["out_noise_bf16"], arrays["grad_noise_bf16"] = out_noise16, grad_noise16
            np.savez_compressed(
                os.path.join(GOLDEN, f"family_{script}_{mode}.npz"), seed=0, input_seed=11, B=2, S=128,
                n_out=len(out), keys=np.array(keys), param_names=np.array([n for n, _ in named]),
                state_digest=np.stack([digest(sd0[k].float()) for k in keys]),
                losses=np.array([l.item() for l in losses], dtype=np.float64),
                grad_is_none=np.array([p.grad is None for _, p in named]),
                grad_digest=np.stack([digest(p.grad) if p.grad is not None else np.zeros(SAMPLE + 2)
                                      for _, p in named]),
                grad_norm=np.array([p.grad.double().norm().item() if p.grad is not None else 0.0 for _, p in named]),
                after_digest=np.stack([digest(net.state_dict()[k].float()) for k in keys]),
                out_noise_fp64=out_noise, grad_noise_fp64=grad_noise, **arrays)


def main():
    if not refload.available():
        raise SystemExit("reference tree not found; goldens can only be generated where /root/reference exists")
    os.makedirs(GOLDEN, exist_ok=True)
    golden_model_s()
    golden_model_c1()
    golden_targets()
    golden_pckh()
    golden_pckh_d()
    golden_aspp()
    golden_families()
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
