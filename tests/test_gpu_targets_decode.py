"""GPU parity tests of target rendering, argmax decode and PCKh against the numpy oracles (oracle/targets_np.py,
oracle/pckh_np.py, both pinned to the reference / Pillow by the `not gpu` tests).  Integer results (label maps,
decoded indices, PCKh counts) must be bit-exact; Gaussians are float64-evaluated and must match to the last
float32 bit except for the documented <= 1 ulp allowance of the device exp()."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

import progressive_process_for_human_pose_estimation_b200 as hg  # noqa: E402
import progressive_process_for_human_pose_estimation_b200.only_one_hourgless as ooh  # noqa: E402
import progressive_process_for_human_pose_estimation_b200.try_with_torch as twt  # noqa: E402
from oracle import pckh_np, targets_np  # noqa: E402


def synth_people(seed, B, P, J, float_xy=False):
    r = np.random.RandomState(seed)
    wh = np.stack([r.randint(200, 700, B), r.randint(200, 700, B)], 1).astype(np.float64)
    kp = np.zeros([B, P, J, 3])
    for b in range(B):
        kp[b, :, :, 0] = r.uniform(0, wh[b, 0], [P, J]) if float_xy else r.randint(0, wh[b, 0], [P, J])
        kp[b, :, :, 1] = r.uniform(0, wh[b, 1], [P, J]) if float_xy else r.randint(0, wh[b, 1], [P, J])
    kp[..., 2] = r.randint(0, 3, [B, P, J])
    npers = r.randint(1, P + 1, B).astype(np.int32)
    return kp, wh, npers


def ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


VARIANTS = [
    dict(truncate=True, accumulate=False),                       # G1 try_with_torch.py:107-132
    dict(truncate=False, accumulate=False, pre_scale=100.0),     # G2 try_with_torch_100.py:64-85
    dict(truncate=False, accumulate=False),                      # G3 only_one_hourgless.py:112-132
    dict(truncate=False, accumulate=True, center_mode=1),        # G4 hourglass_compare.py:713-734
    dict(truncate=True, accumulate=True),                        # G5 hourglass_compare.py:286-313
    dict(truncate=False, accumulate=False, amplitude=1.0 / (2 * np.pi)),  # G6 data_argumentation.py:33-52
]


@pytest.mark.parametrize("variant", VARIANTS)
def test_gaussian_heatmaps_match_numpy(variant):
    kp, wh, npers = synth_people(3, 6, 3, 17, float_xy=not variant["truncate"])
    got = hg.gaussian_heatmaps(kp, wh, num_persons=npers, **variant).cpu().numpy()
    worst = 0
    for b in range(kp.shape[0]):
        ref = targets_np.gauss_map(kp[b, :npers[b]], wh[b], 17, **variant)
        d = ulp_diff(got[b], ref)
        # subnormal / tiny tails: compare absolutely as well
        bad = (d > 1) & (np.abs(got[b] - ref) > 1e-37)
        assert not bad.any(), f"image {b}: {bad.sum()} elements differ by more than 1 ulp"
        worst = max(worst, int(d[np.abs(ref) > 1e-30].max(initial=0)))
    assert worst <= 1
    if variant["truncate"]:
        # integer centres: d2 is an exact integer, every element must be bit-identical in practice
        frac_exact = np.mean([np.array_equal(got[b], targets_np.gauss_map(kp[b, :npers[b]], wh[b], 17, **variant))
                              for b in range(kp.shape[0])])
        assert frac_exact >= 0.5


def test_gaussian_module_api_and_last_person_wins():
    kp, wh, _ = synth_people(5, 4, 2, 17)
    kp[..., 2] = 2
    out = twt.render_targets(kp, wh)
    assert out.shape == (4, 17, 64, 64) and out.dtype == torch.float32
    last_only = twt.render_targets(kp[:, 1:], wh)
    assert torch.equal(out, last_only)  # quirk Q7
    peak = hg.decode_argmax(out)[0].cpu().numpy()
    cx = np.trunc(kp[:, 1, :, 0] / wh[:, None, 0] * 64)
    cy = np.trunc(kp[:, 1, :, 1] / wh[:, None, 1] * 64)
    assert np.array_equal(peak[..., 1], cx.astype(np.int32)) and np.array_equal(peak[..., 0], cy.astype(np.int32))


@pytest.mark.parametrize("mode", ["skeleton", "background", "keypoints", "both"])
def test_label_maps_bit_exact(mode):
    kp, wh, npers = synth_people(11, 8, 3, 17)
    kp[0, :, :, 2] = 0          # image without visible joints -> empty map
    kp[1, 0, 3, 0] = wh[1, 0]   # x == w -> column 64, clipped by the canvas
    kw = dict(skeleton=dict(draw_lines=True), background=dict(draw_lines=True, line_value=1),
              keypoints=dict(draw_points=True, draw_lines=False), both=dict(draw_points=True, draw_lines=True))[mode]
    got = hg.label_maps(kp, wh, twt.sks, num_persons=npers, **kw)
    assert got.dtype == torch.int64 and got.shape == (8, 64, 64)
    for b in range(8):
        ref = targets_np.label_map(kp[b, :npers[b]], wh[b], 17, twt.sks, **kw)
        assert np.array_equal(got[b].cpu().numpy(), ref), f"image {b}"
    assert int(got[0].abs().sum()) == 0


def test_merge_script_targets_bit_exact():
    """try_skeleton_from_keypoints_merge.render_targets: Gaussians (last person) + skeleton map with value = limb index
    (reference :91-135; oracle pinned to the reference dataset class by tests/test_oracle_targets_pckh.py)."""
    import progressive_process_for_human_pose_estimation_b200.try_skeleton_from_keypoints_merge as mg
    kp, wh, npers = synth_people(21, 8, 3, 17)
    gauss, smap = mg.render_targets(kp, wh, num_persons=npers)
    assert gauss.shape == (8, 17, 64, 64) and smap.shape == (8, 64, 64) and smap.dtype == torch.int64
    for b in range(8):
        ref = targets_np.label_map(kp[b, :npers[b]], wh[b], 17, mg.sks, line_value=-1)
        assert np.array_equal(smap[b].cpu().numpy(), ref), b
        gref = targets_np.gauss_map(kp[b, :npers[b]], wh[b], 17, truncate=True)
        assert np.abs(gauss[b].cpu().numpy() - gref).max() <= 1.2e-7


def test_annotation_table_coco_matches_reference_dataset_golden():
    """Annotation -> keypoint tensor -> targets entirely on the device (N1): the COCO annotations of the reference-dataset
    golden (tests/golden/targets_coco.npz, produced by try_different_stack.myImageDataset_COCO through the fake-COCO
    shim) uploaded once as an AnnotationTable; batches of image indices reproduce the reference's Gaussian, skeleton and
    background maps."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "targets_coco.npz"))
    kp, npers, wh = g["keypoints"], g["num_persons"], g["img_wh"]
    persons = [[kp[i, p].reshape(-1).astype(np.int64).tolist() for p in range(npers[i])] for i in range(len(kp))]
    table = hg.AnnotationTable.from_coco(persons, wh)
    assert len(table) == len(kp)
    order = [5, 0, 3, 3, 1, 2, 4]
    k, n, w = table.batch(order)
    assert k.is_cuda and k.dtype == torch.float64 and k.shape[0] == len(order) and k.shape[2:] == (17, 3)
    assert np.array_equal(n.cpu().numpy(), npers[order]) and np.array_equal(w.cpu().numpy(), wh[order])
    for j, i in enumerate(order):
        assert np.array_equal(k[j, :npers[i]].cpu().numpy(), targets_np.coco_persons_to_dense(persons[i]))
    gauss = hg.gaussian_heatmaps(k, w, num_persons=n, truncate=True)
    skel = hg.label_maps(k, w, g["limbs"], num_persons=n)
    bg = hg.label_maps(k, w, g["limbs"], num_persons=n, line_value=1)
    assert ulp_diff(gauss.cpu().numpy(), g["gauss"][order]).max() <= 1
    assert np.array_equal(skel.cpu().numpy(), g["skeleton"][order])
    assert np.array_equal(bg.cpu().numpy(), g["background"][order])
    # an image with more persons than the tensor holds keeps its LAST ones (quirk Q7: the last person wins)
    k1, n1, _ = table.batch(order, max_persons=1)
    assert int(n1.max()) == 1
    g1 = hg.gaussian_heatmaps(k1, w, num_persons=n1, truncate=True)
    assert torch.equal(g1, gauss)


def test_annotation_table_mpii_scatter():
    """MPII: sparse (id, x, y, is_visible) point records -> dense [16, 3] rows (hourglass_compare.py:691-703): missing
    joints stay (0, 0, 0), is_visible == 0 -> invisible, duplicates: the last record wins."""
    r = np.random.RandomState(9)
    samples, sizes = [], []
    for i in range(12):
        ids = r.choice(16, r.randint(0, 17), replace=False).tolist()
        pts = [(j, float(r.uniform(0, 900)), float(r.uniform(0, 700)), int(r.randint(0, 2))) for j in ids]
        if i % 3 == 0 and ids:
            pts.append((ids[0], 11.0, 12.0, 1))   # duplicate id
        samples.append(pts)
        sizes.append((float(r.randint(200, 1000)), float(r.randint(200, 1000))))
    table = hg.AnnotationTable.from_mpii(samples, sizes)
    idx = [3, 0, 11, 7, 0]
    k, n, w = table.batch(idx)
    assert k.shape == (5, 1, 16, 3) and int(n.min()) == 1 and int(n.max()) == 1
    for j, i in enumerate(idx):
        assert np.array_equal(k[j, 0].cpu().numpy(), targets_np.mpii_points_to_dense(samples[i]))
        assert tuple(w[j].cpu().numpy()) == sizes[i]
    gm = hg.gaussian_heatmaps(k, w, truncate=False, accumulate=True, center_mode=1)
    for j, i in enumerate(idx):
        ref = targets_np.gauss_map(targets_np.mpii_points_to_dense(samples[i])[None], sizes[i], 16, center_mode=1,
                                   truncate=False, accumulate=True)
        assert ulp_diff(gm[j].cpu().numpy(), ref).max() <= 1


def test_mpii_label_maps_bit_exact():
    """MPII keypoint (ImageDraw.ellipse on the float centre) and skeleton maps of train.py:668-690 against the numpy
    oracle (pinned to Pillow by tests/test_oracle_targets_pckh.py::test_mpii_label_maps_match_pillow)."""
    sks = [[0, 1], [1, 2], [2, 6], [6, 3], [3, 4], [4, 5], [6, 7], [7, 8], [8, 9], [10, 11], [11, 12], [12, 7], [7, 13],
           [13, 14], [14, 15]]
    r = np.random.RandomState(4)
    B = 16
    wh = np.stack([r.randint(150, 900, B), r.randint(150, 900, B)], 1).astype(np.float64)
    kp = np.zeros([B, 1, 16, 3])
    kp[:, 0, :, 0] = r.uniform(-6, 1, [B, 16]) + r.rand(B, 16) * wh[:, :1]
    kp[:, 0, :, 1] = r.uniform(-6, 1, [B, 16]) + r.rand(B, 16) * wh[:, 1:]
    kp[0, 0, :4, 0] = np.array([0.1, 0.49, 0.5, 63.6]) * wh[0, 0] / 64      # collapsed / clipped ellipse boxes
    kp[0, 0, :4, 1] = np.array([0.2, 0.3, 63.9, 0.4]) * wh[0, 1] / 64
    kp[..., 2] = r.rand(B, 1, 16) < 0.8
    gk = hg.label_maps(kp, wh, sks, center_mode=1, draw_points="ellipse", draw_lines=False)
    gs = hg.label_maps(kp, wh, sks, center_mode=1, draw_lines=True)
    for b in range(B):
        rk = targets_np.label_map(kp[b], wh[b], 16, sks, center_mode=1, draw_points=2, draw_lines=False)
        rs = targets_np.label_map(kp[b], wh[b], 16, sks, center_mode=1, draw_lines=True)
        assert np.array_equal(gk[b].cpu().numpy(), rk), b
        assert np.array_equal(gs[b].cpu().numpy(), rs), b
    assert int((gk > 0).sum()) > 0


def _heatmaps(seed, B, C, dtype):
    r = np.random.RandomState(seed)
    x = torch.from_numpy(r.randn(B, C, 64, 64).astype(np.float32))
    x[0, 0] = 0.0                           # constant map -> index 0
    x[1, 3, 10, 5] = 9.0
    x[1, 3, 10, 7] = 9.0                    # duplicated maximum -> lowest index
    x[2, 1] = x[2, 1].half().float()        # fp16-quantised maps have many ties
    x[3, 2, 63, 63] = 50.0                  # last element
    return x.to(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_decode_argmax_bit_exact(dtype):
    x = _heatmaps(0, 4, 16, dtype)
    yx, mx = hg.decode_argmax(x.cuda())
    ref = np.array([[pckh_np.argmax_first(x[b, c].float().numpy()) for c in range(16)] for b in range(4)])
    assert np.array_equal(yx.cpu().numpy(), ref)
    assert torch.equal(mx.cpu(), x.float().amax((2, 3)))
    # first index of torch.nonzero(x >= max), the reference's own expression (hourglass_compare.py:831)
    for b, c in [(0, 0), (1, 3), (2, 1), (3, 2)]:
        hm = x[b, c].float()
        assert tuple(torch.nonzero(hm >= hm.max())[0].tolist()) == tuple(yx[b, c].tolist())


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("chan_offset", [0, 1])
def test_pckh_sweep_bit_exact(dtype, chan_offset):
    for seed in range(5):
        r = np.random.RandomState(100 + seed)
        B, J = 6, 16
        x = _heatmaps(seed, B, J + chan_offset, dtype)
        tgt = torch.zeros(B, 64, 64, dtype=torch.long)
        for b in range(B):
            for j in range(J):
                if b != 4 and r.rand() < 0.85:   # image 4: no joints at all -> NaN accuracy row
                    tgt[b, r.randint(64), r.randint(64)] = j + 1
        tgt[5, 0, 0] = 3
        tgt[5, 40, 2] = 3                        # a joint labelled twice -> first row-major position
        rect = torch.from_numpy(r.uniform(0, 64, size=(B, 4)).astype(np.float32))
        got = hg.pckh_sweep_counts(x.cuda(), tgt.cuda(), rect.cuda(), chan_offset)
        ref = pckh_np.pckh_sweep(x.float().numpy(), tgt.numpy(), rect.numpy(), chan_offset)
        for k in ("correct", "total", "predict", "label", "found"):
            assert np.array_equal(got[k].cpu().numpy(), ref[k]), (seed, k)
        assert np.array_equal(got["standard"].cpu().numpy(), ref["standard"])
        mod = hg.PCKh_hourglass() if chan_offset == 0 else hg.PCKh_softmax()
        res = mod(x.cuda(), tgt.cuda(), rect.cuda())
        acc = res[0]
        assert acc.shape == (B, 11) and acc.dtype == np.float64
        assert np.array_equal(np.isnan(acc), np.isnan(ref["accuracy"]))
        assert np.array_equal(np.nan_to_num(acc), np.nan_to_num(ref["accuracy"]))
        assert np.isnan(acc[4]).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_pckh_d_bit_exact(dtype):
    """PCKh D (calculate_parameters.py:906-937) against the numpy oracle and the reference-generated golden."""
    from oracle.synth import pckh_inputs, pckh_near_inputs

    for seed in range(5):
        d = pckh_near_inputs(seed)
        x = torch.from_numpy(d["x17"]).to(dtype)
        tgt, rect = torch.from_numpy(d["target"]), torch.from_numpy(d["rect"])
        acc, pred, lab = hg.PCKh_half_standard()(x.cuda(), tgt.cuda(), rect.cuda())
        c, t, p, l = pckh_np.pckh_d(x.float().numpy(), d["target"], d["rect"])
        assert acc == [int(a) / int(b) for a, b in zip(c, t)], seed
        assert np.array_equal(np.stack(pred), p) and np.array_equal(np.stack(lab), l)
        if seed == 0 and dtype == torch.float32:
            g = np.load(os.path.join(GOLDEN, "pckh_d.npz"))
            assert np.array_equal(np.array(acc), g["acc_d"])
            assert np.array_equal(np.stack(pred), g["pred_d"]) and np.array_equal(np.stack(lab), g["lab_d"])
    d = pckh_inputs(0)  # image 4 has no annotated joint: correct / total is 0 / 0 in the reference too
    with pytest.raises(ZeroDivisionError):
        hg.PCKh_half_standard()(torch.from_numpy(d["x17"]).cuda(), torch.from_numpy(d["target"]).cuda(),
                                torch.from_numpy(d["rect"]).cuda())


def test_pckh_a_bit_exact_counts():
    for seed in range(4):
        r = np.random.RandomState(seed)
        B = 5
        ooh.batch_size = B
        tgt = torch.from_numpy(r.rand(B, 14, 64, 64).astype(np.float32))
        tgt[0, 2] = 0                           # absent joint -> skipped
        tgt[1, 13, 5, 5] = 3.0
        tgt[1, 13, 5, 9] = 3.0                  # tie in the head map
        x = torch.from_numpy(r.rand(B, 18, 64, 64).astype(np.float32))
        x[:, :14] = 0.7 * tgt + 0.3 * x[:, :14]
        pck = ooh.PCKh()
        c = pck.counts(x.cuda(), tgt.cuda()).cpu().numpy()
        rc, rt = pckh_np.pckh_a(x.numpy(), tgt.numpy(), B)
        assert (int(c[0]), int(c[1])) == (rc, rt)
        assert pck(x.cuda(), tgt.cuda()) == rc / rt
    ooh.batch_size = 16


def test_render_decode_roundtrip_at_scale():
    """Full-size property (BASELINE config 5 batch): decode(render(kp)) returns the truncated centres."""
    B, J = 256, 17
    r = np.random.RandomState(7)
    wh = np.tile(np.array([[640.0, 480.0]]), (B, 1))
    kp = np.zeros([B, 1, J, 3])
    kp[..., 0] = r.randint(0, 640, [B, 1, J])
    kp[..., 1] = r.randint(0, 480, [B, 1, J])
    kp[..., 2] = 2
    hm = hg.gaussian_heatmaps(kp, wh)
    yx, mx = hg.decode_argmax(hm)
    assert torch.all(mx == 1.0)
    cx = np.trunc(kp[:, 0, :, 0] / 640.0 * 64).astype(np.int32)
    cy = np.trunc(kp[:, 0, :, 1] / 480.0 * 64).astype(np.int32)
    assert np.array_equal(yx[..., 1].cpu().numpy(), cx) and np.array_equal(yx[..., 0].cpu().numpy(), cy)


def test_decode_requires_cuda_tensor():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hg.decode_argmax(torch.zeros(1, 1, 64, 64))


def test_fused_mse_losses_match_stock_modules():
    """hg.mse_losses == sum of nn.MSELoss per stack (try_with_torch.py:333-341): values and gradients."""
    torch.manual_seed(0)
    S, shape = 8, (4, 16, 64, 64)
    tgt = torch.rand(shape, device="cuda")
    preds = [torch.randn(shape, device="cuda", requires_grad=True) for _ in range(S)]
    ref = [torch.nn.MSELoss()(p, tgt) for p in preds]
    sum(w * l for w, l in zip(range(1, S + 1), ref)).backward()
    gref = [p.grad.clone() for p in preds]
    for p in preds:
        p.grad = None
    losses = hg.mse_losses(preds, tgt)
    assert losses.shape == (S,)
    assert torch.allclose(losses, torch.stack(ref).detach(), rtol=1e-5, atol=0)
    (losses * torch.arange(1, S + 1, device="cuda")).sum().backward()
    for p, g in zip(preds, gref):
        assert torch.allclose(p.grad, g, rtol=1e-6, atol=1e-9)
    # odd element count (scalar tail) and a prediction that needs no gradient
    a = torch.randn(3, 5, 7, device="cuda", requires_grad=True)
    b = torch.randn(3, 5, 7, device="cuda")
    t = torch.rand(3, 5, 7, device="cuda")
    l2 = hg.mse_losses([a, b], t)
    assert torch.allclose(l2, torch.stack([((a - t) ** 2).mean(), ((b - t) ** 2).mean()]).detach(), rtol=1e-5)
    l2.sum().backward()
    assert torch.allclose(a.grad, 2 * (a.detach() - t) / t.numel(), rtol=1e-6, atol=1e-9)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hg.mse_losses([a.detach().cpu()], t.cpu())


def test_fused_cross_entropy_heads_match_stock_modules():
    """hg.cross_entropy_losses == the nn.CrossEntropyLoss heads of try_skeleton_and_keypoints.py:423-435 (channel
    slices [:, :18] / [:, 18:] of 4 stack outputs), only_one_hourgless.py:370 and try_different_stack.py:388-389,
    evaluated by the stock module on the CPU in fp32 (the reference's implementation); rtol 1e-5 (north_star fp32)."""
    def close(got, want):
        # rtol 1e-5 plus an absolute term at 1e-5 of the gradient scale: softmax - onehot cancels near sm = 1
        return torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5 * want.abs().max().item())

    torch.manual_seed(0)
    B, S = 4, 4
    outs = [(3 * torch.randn(B, 38, 64, 64)).requires_grad_() for _ in range(S)]
    yk = torch.randint(0, 18, (B, 64, 64))
    ys = torch.randint(0, 20, (B, 64, 64))
    ce = torch.nn.CrossEntropyLoss()
    ref = [ce(o[:, :18], yk) for o in outs] + [ce(o[:, 18:], ys) for o in outs]
    w = torch.arange(1, 2 * S + 1, dtype=torch.float32)
    sum(wi * li for wi, li in zip(w, ref)).backward()
    gref = [o.grad.clone() for o in outs]
    dev = [o.detach().cuda().requires_grad_() for o in outs]
    terms = [(o, yk.cuda(), (0, 18)) for o in dev] + [(o, ys.cuda(), (18, 38)) for o in dev]
    n0 = hg._lib.launch_count()
    losses = hg.cross_entropy_losses(terms)
    assert hg._lib.launch_count() - n0 == 2  # label count + fused loss/gradient: 8 terms in 2 launches
    assert torch.allclose(losses.cpu(), torch.stack(ref).detach(), rtol=1e-5, atol=0)
    (losses * w.cuda()).sum().backward()
    for o, g in zip(dev, gref):
        assert close(o.grad, g)
    # the same heads called on channel-slice VIEWS (how the reference script passes them), no explicit slices
    dev2 = [o.detach().cuda().requires_grad_() for o in outs]
    l2 = hg.cross_entropy_losses([(o[:, :18], yk.cuda()) for o in dev2] + [(o[:, 18:], ys.cuda()) for o in dev2])
    assert torch.allclose(l2, losses, rtol=1e-6)  # block partial sums are added with atomics: order varies
    (l2 * w.cuda()).sum().backward()
    for o, g in zip(dev2, gref):
        assert close(o.grad, g)
    # single 18-class head (only_one_hourgless.py:370); a partial slice leaves the other channels' gradient at zero
    x = torch.randn(2, 18, 64, 64, requires_grad=True)
    y = torch.randint(0, 18, (2, 64, 64))
    r = ce(x, y)
    r.backward()
    xd = x.detach().cuda().requires_grad_()
    l = hg.cross_entropy_losses([(xd, y.cuda())])
    l.sum().backward()
    assert torch.allclose(l.cpu()[0], r.detach(), rtol=1e-5) and close(xd.grad, x.grad)
    xd2 = x.detach().cuda().requires_grad_()
    hg.cross_entropy_losses([(xd2, (y % 5).cuda(), (3, 8))]).sum().backward()
    assert (xd2.grad[:, :3] == 0).all() and (xd2.grad[:, 8:] == 0).all() and (xd2.grad[:, 3:8] != 0).any()
    # ignore_index (module default -100): mean over the remaining labels; all ignored -> NaN like PyTorch
    yi = y.clone()
    yi[0, :32] = -100
    x.grad = None
    ri = ce(x, yi)
    ri.backward()
    xd3 = x.detach().cuda().requires_grad_()
    li = hg.cross_entropy_losses([(xd3, yi.cuda())])
    li.sum().backward()
    assert torch.allclose(li.cpu()[0], ri.detach(), rtol=1e-5)
    assert close(xd3.grad, x.grad) and (xd3.grad[0, :, :32] == 0).all()
    assert torch.isnan(hg.cross_entropy_losses([(xd3.detach(), torch.full_like(yi, -100).cuda())])).all()
    with pytest.raises(IndexError):
        hg.cross_entropy_losses([(xd3.detach(), (y + 1).cuda())], check_labels=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hg.cross_entropy_losses([(x.detach(), y)])


def test_to_tensor_normalize_bit_exact():
    """hg.to_tensor_normalize == transforms.ToTensor() + Normalize(0.5, 0.5) of try_with_torch.py:310-313 evaluated
    on the CPU (torchvision when importable, else the same two torch expressions), every uint8 value, bit for bit."""
    r = np.random.RandomState(0)
    img = r.randint(0, 256, size=(5, 64, 48, 3)).astype(np.uint8)
    img[0, 0, :, 0] = np.arange(48) * 5 % 256
    img[1].reshape(-1)[:256] = np.arange(256)          # all 256 values
    for mean, std in (((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        try:
            from PIL import Image
            from torchvision import transforms
            tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=mean, std=std)])
            ref = torch.stack([tf(Image.fromarray(im)) for im in img])
        except ImportError:
            t = torch.from_numpy(img).permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255)
            ref = t.sub_(torch.tensor(mean).view(1, 3, 1, 1)).div_(torch.tensor(std).view(1, 3, 1, 1))
        got = hg.to_tensor_normalize(torch.from_numpy(img).cuda(), mean, std)
        assert got.shape == (5, 3, 64, 48) and got.dtype == torch.float32
        assert torch.equal(got.cpu(), ref)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hg.to_tensor_normalize(torch.from_numpy(img))


def test_pckh_from_logits_equals_softmax_then_pckh():
    """Fused softmax -> PCKh B (hourglass_compare.py:1160: pckh.forward(softmax(result[2]), y, rect)): identical
    counts and decoded positions to (a) PyTorch's softmax on the GPU followed by the un-fused evaluator and (b) the
    numpy oracle on the CPU softmax, on peaked and on random class scores."""
    from oracle.synth import pckh_near_inputs

    for seed in range(4):
        d = pckh_near_inputs(seed)
        r = np.random.RandomState(50 + seed)
        z = r.randn(*d["x17"].shape).astype(np.float32) * 2
        if seed % 2 == 0:   # peaked scores: the labelled joints' neighbourhoods win
            z += 6 * np.log(np.maximum(d["x17"], 1e-6)).astype(np.float32).clip(-3, 0)
        zt = torch.from_numpy(z)
        tgt, rect = torch.from_numpy(d["target"]), torch.from_numpy(d["rect"])
        fused = hg.pckh_sweep_counts(zt.cuda(), tgt.cuda(), rect.cuda(), 1, logits=True)
        plain = hg.pckh_sweep_counts(torch.softmax(zt.cuda(), 1), tgt.cuda(), rect.cuda(), 1)
        for k in ("correct", "total", "predict", "label", "found"):
            assert torch.equal(fused[k], plain[k]), (seed, k)
        ref = pckh_np.pckh_sweep(torch.softmax(zt, 1).numpy(), d["target"], d["rect"], 1)
        for k in ("correct", "total", "predict", "label", "found"):
            assert np.array_equal(fused[k].cpu().numpy(), ref[k]), (seed, k)
        acc, pred, lab, std = hg.PCKh_from_logits()(zt.cuda(), tgt.cuda(), rect.cuda())
        assert np.array_equal(acc, ref["accuracy"]) and np.array_equal(np.stack(pred), ref["predict"].astype(np.float64))
    with pytest.raises(RuntimeError, match="fp32"):
        hg.pckh_sweep_counts(zt.cuda().half(), tgt.cuda(), rect.cuda(), 1, logits=True)


def test_resize_bicubic_bit_exact_with_pillow():
    """hg.resize_bicubic == [Image.resize([256, 256]) ...] of try_with_torch.py:99 run by Pillow itself on the host, for a
    ragged batch; the fused variant == transforms.ToTensor + Normalize(0.5, 0.5) of the resized image (:310-313)."""
    from PIL import Image
    r = np.random.RandomState(1)
    sizes = [(480, 640), (427, 640), (640, 480), (256, 256), (100, 37), (333, 500), (64, 300), (257, 255), (3, 5), (1, 1),
             (720, 1280)]
    imgs = []
    for h, w in sizes:
        im = r.randint(0, 256, (h, w, 3)).astype(np.uint8)
        if h > 300:
            im[: h // 2] = np.linspace(0, 255, w)[None, :, None].astype(np.uint8)
        imgs.append(im)
    want = np.stack([np.array(Image.fromarray(im).convert("RGB").resize([256, 256])) for im in imgs])
    n0 = hg._lib.launch_count()
    got = hg.resize_bicubic([torch.from_numpy(im).cuda() for im in imgs])
    assert hg._lib.launch_count() - n0 == 2          # horizontal + vertical pass for the whole ragged batch
    assert got.dtype == torch.uint8 and got.shape == (len(sizes), 256, 256, 3)
    assert np.array_equal(got.cpu().numpy(), want)
    fused = hg.resize_bicubic([torch.from_numpy(im).cuda() for im in imgs], normalize=((0.5,) * 3, (0.5,) * 3))
    ref = torch.from_numpy(want).permute(0, 3, 1, 2).contiguous().to(torch.float32).div(255).sub_(0.5).div_(0.5)
    assert torch.equal(fused.cpu(), ref)
    small = hg.resize_bicubic([torch.from_numpy(imgs[0]).cuda()], size=(64, 128))
    assert np.array_equal(small[0].cpu().numpy(), np.array(Image.fromarray(imgs[0]).resize([64, 128])))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hg.resize_bicubic([torch.from_numpy(imgs[0])])
