"""`not gpu`: pins the numpy oracles of target rendering and PCKh to the golden vectors produced by the real
reference (tests/golden/targets_coco.npz, pckh.npz), to Pillow's own rasteriser, and to the reference classes
when /root/reference exists."""
import os

import numpy as np
import pytest
import torch

from oracle import pckh_np, refload, targets_np
from oracle.synth import pckh_inputs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_gauss_and_label_maps_reproduce_reference_dataset_golden():
    g = np.load(os.path.join(GOLDEN, "targets_coco.npz"))
    kp, npers, wh, limbs = g["keypoints"], g["num_persons"], g["img_wh"], g["limbs"].tolist()
    for i in range(kp.shape[0]):
        people = kp[i, :npers[i]]
        # try_different_stack.py:121-155 (Gaussian: last person wins; skeleton value i+1; background value 1)
        np.testing.assert_array_equal(targets_np.gauss_map(people, wh[i], 17, truncate=True), g["gauss"][i])
        np.testing.assert_array_equal(targets_np.label_map(people, wh[i], 17, limbs), g["skeleton"][i])
        np.testing.assert_array_equal(targets_np.label_map(people, wh[i], 17, limbs, line_value=1), g["background"][i])
        # try_skeleton_and_keypoints.py:93-114 (keypoint points value k+1; skeleton lines)
        np.testing.assert_array_equal(
            targets_np.label_map(people, wh[i], 17, limbs, draw_points=True, draw_lines=False), g["keypoint_map"][i])
        np.testing.assert_array_equal(targets_np.label_map(people, wh[i], 17, limbs), g["skeleton2"][i])


def test_rasteriser_matches_pillow():
    from PIL import Image, ImageDraw

    rng = np.random.RandomState(0)
    for it in range(4000):
        c = rng.uniform(-10, 74, size=4)
        if it % 3 == 0:
            c = np.trunc(c)
        if it % 7 == 0:
            c[2:] = c[:2]  # degenerate line = one point
        im = Image.fromarray(np.zeros([64, 64]), "L")
        ImageDraw.Draw(im).line(c.tolist(), "rgb(7, 7, 7)")
        cv = np.zeros([64, 64], np.uint8)
        targets_np.draw_line(cv, *c, 7)
        assert np.array_equal(np.array(im), cv), c
    for it in range(500):
        c = rng.uniform(-3, 67, size=2)
        im = Image.fromarray(np.zeros([64, 64]), "L")
        ImageDraw.Draw(im).point(c.tolist(), "rgb(5, 5, 5)")
        cv = np.zeros([64, 64], np.uint8)
        targets_np.draw_point(cv, *c, 5)
        assert np.array_equal(np.array(im), cv), c


def test_mpii_label_maps_match_pillow():
    """MPII keypoint / skeleton label maps (train.py:668-690): the oracle against the reference's own statements run
    with Pillow (the third-party rasteriser of this path; 12.2.0 in this image): ImageDraw.ellipse on the float
    centre +-0.5 and ImageDraw.line on float end points, including centres on and beyond the canvas border."""
    from PIL import Image, ImageDraw
    sks = [[0, 1], [1, 2], [2, 6], [6, 3], [3, 4], [4, 5], [6, 7], [7, 8], [8, 9], [10, 11], [11, 12], [12, 7], [7, 13],
           [13, 14], [14, 15]]
    r = np.random.RandomState(3)
    for trial in range(40):
        w, h = r.randint(150, 900), r.randint(150, 900)
        pts = np.zeros([16, 3])
        pts[:, 0] = r.uniform(-6, w + 6, 16)
        pts[:, 1] = r.uniform(-6, h + 6, 16)
        if trial % 5 == 0:
            pts[:4, 0] = [0.1 * w / 64, 0.49 * w / 64, 0.5 * w / 64, 63.6 * w / 64]   # box collapses / clips
            pts[:4, 1] = [0.2 * h / 64, 0.3 * h / 64, 63.9 * h / 64, 0.4 * h / 64]
        pts[:, 2] = r.rand(16) < 0.8
        inputsize = 256
        kmap = Image.fromarray(np.zeros([64, 64], dtype=np.uint8))
        smap = Image.fromarray(np.zeros([64, 64], dtype=np.uint8))
        dk, ds = ImageDraw.Draw(kmap), ImageDraw.Draw(smap)
        xs = pts[:, 0] * inputsize / w / 4
        ys = pts[:, 1] * inputsize / h / 4
        v = pts[:, 2]
        for i in range(16):
            if v[i] > 0:
                size = 1
                dk.ellipse((xs[i] - size / 2, ys[i] - size / 2, xs[i] + size / 2, ys[i] + size / 2), fill=i + 1)
        for i, sk in enumerate(sks):
            if np.all(v[sk]) > 0:
                ds.line(np.stack([xs[sk], ys[sk]], axis=1).reshape([-1]).tolist(), i + 1)
        got_k = targets_np.label_map(pts[None], (w, h), 16, sks, center_mode=1, draw_points=2, draw_lines=False)
        got_s = targets_np.label_map(pts[None], (w, h), 16, sks, center_mode=1, draw_points=False, draw_lines=True)
        assert np.array_equal(got_k, np.array(kmap).astype(np.int64)), trial
        assert np.array_equal(got_s, np.array(smap).astype(np.int64)), trial


def test_bicubic_tables_reproduce_pillow_resize():
    """`image.resize([256, 256])` of the reference datasets (try_with_torch.py:99): the product's coefficient tables
    (precompute_coeffs + normalize_coeffs_8bpc restated on the host) + the integer passes == Pillow, bit for bit, for
    down-scaling, up-scaling, odd and degenerate sizes."""
    from PIL import Image

    from oracle.resize_np import resize_reference_numpy
    r = np.random.RandomState(0)
    for (h, w) in [(480, 640), (427, 640), (256, 256), (100, 37), (500, 333), (64, 300), (257, 255), (3, 5), (1, 1)]:
        img = r.randint(0, 256, (h, w, 3)).astype(np.uint8)
        if h > 300:
            img[: h // 2] = np.linspace(0, 255, w)[None, :, None].astype(np.uint8)
        want = np.array(Image.fromarray(img).convert("RGB").resize([256, 256]))
        assert np.array_equal(resize_reference_numpy(img), want), (h, w)
    img = r.randint(0, 256, (120, 90, 3)).astype(np.uint8)
    want = np.array(Image.fromarray(img).resize([64, 128]))     # (width, height) like PIL
    assert np.array_equal(resize_reference_numpy(img, 64, 128), want)


def test_gauss_variants_against_reference_expressions():
    """Float-centre / x100 / accumulate variants evaluated with the reference's own numpy expressions."""
    import numpy.matlib  # noqa: F401

    r = np.random.RandomState(1)
    w, h = 500.0, 375.0
    kp = np.stack([r.uniform(0, w, 17), r.uniform(0, h, 17), np.full(17, 2.0)], 1)

    def ref_map(xs, ys, scale):  # try_with_torch_100.py:69-83 verbatim structure
        mask_x = np.matlib.repmat(xs, 64, 64)
        mask_y = np.matlib.repmat(ys, 64, 64)
        x_map = np.matlib.repmat(np.arange(64), 64, 1)
        y_map = np.transpose(np.matlib.repmat(np.arange(64), 64, 1))
        temp = scale * ((x_map - mask_x) ** 2 + (y_map - mask_y) ** 2) / (2 * 1 ** 2)
        return np.exp(-temp)

    want100 = np.stack([ref_map(kp[k, 0] / w * 64, kp[k, 1] / h * 64, 100) for k in range(17)])
    got100 = targets_np.gauss_map(kp[None], (w, h), 17, truncate=False, pre_scale=100.0)
    np.testing.assert_array_equal(got100, torch.Tensor(want100).numpy())
    want_mpii = np.stack([ref_map(kp[k, 0] * 256 / w / 4, kp[k, 1] * 256 / h / 4, 1) for k in range(17)])
    got_mpii = targets_np.gauss_map(np.stack([kp, kp]), (w, h), 17, truncate=False, accumulate=True, center_mode=1)
    np.testing.assert_array_equal(got_mpii, torch.Tensor(want_mpii + want_mpii).numpy())


def test_pckh_oracle_reproduces_reference_golden():
    g = np.load(os.path.join(GOLDEN, "pckh.npz"))
    d = pckh_inputs(int(g["seed"]))
    c = pckh_np.pckh_sweep(d["x"], d["target"], d["rect"], 0)
    np.testing.assert_array_equal(np.nan_to_num(c["accuracy"], nan=-1), np.nan_to_num(g["acc_c"], nan=-1))
    np.testing.assert_array_equal(c["predict"].astype(np.float64), g["pred_c"])
    np.testing.assert_array_equal(c["label"].astype(np.float64), g["lab_c"])
    b = pckh_np.pckh_sweep(d["x17"], d["target"], d["rect"], 1)
    np.testing.assert_array_equal(np.nan_to_num(b["accuracy"], nan=-1), np.nan_to_num(g["acc_b"], nan=-1))
    np.testing.assert_array_equal(b["predict"].astype(np.float64), g["pred_b"])
    np.testing.assert_array_equal(b["standard"], g["std_b"])
    ca, ta = pckh_np.pckh_a(d["x14"], d["t14"], d["x14"].shape[0])
    assert ca / ta == float(g["acc_a"])
    assert np.isnan(c["accuracy"][4]).all()  # image without annotated joints


def test_pckh_d_oracle_reproduces_reference_golden():
    from oracle.synth import pckh_near_inputs

    g = np.load(os.path.join(GOLDEN, "pckh_d.npz"))
    d = pckh_near_inputs(int(g["seed"]))
    c, t, pred, lab = pckh_np.pckh_d(d["x17"], d["target"], d["rect"])
    np.testing.assert_array_equal(c.astype(np.float64) / t, g["acc_d"])
    assert 0 < c.sum() < t.sum()  # both outcomes of the threshold test occur
    np.testing.assert_array_equal(pred, g["pred_d"])
    np.testing.assert_array_equal(lab, g["lab_d"])
    b = pckh_np.pckh_sweep(d["x17"], d["target"], d["rect"], 1)
    np.testing.assert_array_equal(b["accuracy"], g["acc_b"])
    np.testing.assert_array_equal(b["predict"].astype(np.float64), g["pred_b"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_pckh_d_oracle_vs_reference_class_many_seeds():
    from oracle.synth import pckh_near_inputs

    cp = refload.load("calculate_parameters")
    for seed in range(1, 5):
        d = pckh_near_inputs(seed)
        acc, pred, lab = cp.PCKh().forward(torch.from_numpy(d["x17"]), torch.from_numpy(d["target"]),
                                           torch.from_numpy(d["rect"]))
        c, t, p, l = pckh_np.pckh_d(d["x17"], d["target"], d["rect"])
        assert acc == [int(a) / int(b) for a, b in zip(c, t)]
        assert np.array_equal(np.stack(pred), p) and np.array_equal(np.stack(lab), l)
    d = pckh_inputs(0)  # image 4 has no annotated joint: the reference divides 0 / 0
    with pytest.raises(ZeroDivisionError):
        cp.PCKh().forward(torch.from_numpy(d["x17"]), torch.from_numpy(d["target"]), torch.from_numpy(d["rect"]))


def test_thresholds_are_float32_rounded():
    want = [0, 0.05000000075, 0.10000000149, 0.15000000596, 0.20000000298, 0.25, 0.30000001192, 0.34999999404,
            0.40000000596, 0.44999998808, 0.5]
    np.testing.assert_allclose(pckh_np.THRESHOLDS_F32.astype(np.float64), want, rtol=0, atol=1e-11)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_pckh_oracle_vs_reference_classes_many_seeds():
    hc = refload.load("hourglass_compare")
    pc = refload.load("performance_compare")
    oo = refload.load("only_one_hourgless")
    pc.nKeypoint_MPII = 16
    for seed in range(1, 6):
        d = pckh_inputs(seed)
        tgt, rect = torch.from_numpy(d["target"]), torch.from_numpy(d["rect"])
        acc, pred, lab = hc.PCKh().forward(torch.from_numpy(d["x"]), tgt, rect)
        o = pckh_np.pckh_sweep(d["x"], d["target"], d["rect"])
        assert np.array_equal(np.nan_to_num(acc, nan=-1), np.nan_to_num(o["accuracy"], nan=-1))
        assert np.array_equal(np.stack(pred), o["predict"].astype(np.float64))
        acc2, pred2, lab2, std2 = pc.PCKh().forward(torch.from_numpy(d["x17"]), tgt, rect)
        o2 = pckh_np.pckh_sweep(d["x17"], d["target"], d["rect"], 1)
        assert np.array_equal(np.nan_to_num(acc2, nan=-1), np.nan_to_num(o2["accuracy"], nan=-1))
        assert np.array_equal(np.array([float(s) for s in std2], dtype=np.float32), o2["standard"])
        oo.batch_size = d["x14"].shape[0]
        a = oo.PCKh().forward(torch.from_numpy(d["x14"]), torch.from_numpy(d["t14"]))
        c, t = pckh_np.pckh_a(d["x14"], d["t14"], d["x14"].shape[0])
        assert a == c / t


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_merge_dataset_targets_vs_reference_class():
    """try_skeleton_from_keypoints_merge.myImageDataset_COCO.__getitem__ (reference :91-135) through the fake-COCO
    shim: Gaussians of the last person, skeleton label map drawn with value = limb index over all persons."""
    import tempfile

    from PIL import Image

    from oracle.make_golden import FakeCOCO
    mg = refload.load("try_skeleton_from_keypoints_merge")
    r = np.random.RandomState(2)
    W, H, n_img, P, J = 640, 480, 5, 3, 17
    kp = np.zeros([n_img, P, J, 3])
    kp[..., 0], kp[..., 1] = r.randint(0, W, [n_img, P, J]), r.randint(0, H, [n_img, P, J])
    kp[..., 2] = r.randint(0, 3, [n_img, P, J])
    npers = r.randint(1, P + 1, n_img)
    FakeCOCO.skeleton = (np.array(mg.sks) + 1).tolist()
    FakeCOCO.persons = {i: [kp[i, p].reshape(-1).astype(np.int64).tolist() for p in range(npers[i])] for i in range(n_img)}
    tmp = tempfile.mkdtemp()
    Image.fromarray(np.zeros([H, W, 3], dtype=np.uint8)).save(os.path.join(tmp, "img.jpg"))
    mg.COCO = FakeCOCO
    ds = mg.myImageDataset_COCO("x", tmp, lambda im: torch.zeros(1))
    for i in range(n_img):
        _, g, s = ds[i]
        persons = kp[i, :npers[i]]
        np.testing.assert_array_equal(targets_np.gauss_map(persons, (W, H), J, truncate=True), g.numpy())
        want = targets_np.label_map(persons, (W, H), J, mg.sks, line_value=-1)
        assert np.array_equal(want, s.numpy()), i
