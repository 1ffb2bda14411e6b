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


def main():
    if not refload.available():
        raise SystemExit("reference tree not found; goldens can only be generated where /root/reference exists")
    os.makedirs(GOLDEN, exist_ok=True)
    golden_model_s()
    golden_model_c1()
    golden_targets()
    golden_pckh()
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
