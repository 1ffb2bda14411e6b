"""`not gpu`: pins the CPU oracle of the network (oracle/hourglass_torch.py) to the golden vectors generated from
the real reference (tests/golden/model_*.npz, oracle/make_golden.py) and, where /root/reference exists, to the
reference classes themselves, bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import hourglass_torch as ho
from oracle import refload
from oracle.make_golden import digest, model_inputs

import progressive_process_for_human_pose_estimation_b200.only_one_hourgless as ooh
import progressive_process_for_human_pose_estimation_b200.try_with_torch as twt

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _restore():
    yield
    twt.nStack, twt.nOutChannels = 4, 17
    ooh.nStack, ooh.nOutChannels = 1, 18


def test_oracle_reproduces_golden_train_step():
    """The drop-in's seeded construction gives the reference's weights; the oracle port reproduces the reference's
    forward, losses, gradients and BatchNorm buffer updates stored in the golden file."""
    g = np.load(os.path.join(GOLDEN, "model_s_2stack.npz"))
    twt.nStack, twt.nOutChannels = int(g["nStack"]), int(g["J"])
    torch.manual_seed(int(g["seed"]))
    net = twt.creatModel()  # parameter containers only: no GPU needed to construct
    sd0 = net.state_dict()
    assert list(sd0.keys()) == list(g["keys"])
    assert [n for n, _ in net.named_parameters()] == list(g["param_names"])
    np.testing.assert_array_equal(np.stack([digest(sd0[k].float()) for k in sd0]), g["state_digest"])
    x, tgt = model_inputs(int(g["input_seed"]), int(g["B"]), int(g["S"]), int(g["J"]))
    sd = ho.clone_state(sd0, requires_grad=True)
    out = ho.creat_model_s(sd, x, ho.Config(nStack=2, nOutChannels=16))
    total, per = ho.mse_losses(out, tgt)
    total.backward()
    np.testing.assert_array_equal(out[0].detach().numpy(), g["out0"])
    np.testing.assert_array_equal(out[1].detach().numpy(), g["out1"])
    np.testing.assert_array_equal(np.array([l.item() for l in per]), g["losses"])
    for i, name in enumerate(g["param_names"]):
        gr = sd[str(name)].grad
        assert (gr is None) == bool(g["grad_is_none"][i]), name
        if gr is not None:
            np.testing.assert_array_equal(digest(gr), g["grad_digest"][i], err_msg=str(name))
    np.testing.assert_array_equal(np.stack([digest(sd[k].detach().float()) for k in sd0]), g["after_digest"])


def test_oracle_reproduces_golden_config1_forward():
    g = np.load(os.path.join(GOLDEN, "model_c1_1stack.npz"))
    ooh.nOutChannels = 16
    torch.manual_seed(int(g["seed"]))
    net = ooh.creatModel()
    x, tgt = model_inputs(int(g["input_seed"]), int(g["B"]), int(g["S"]), int(g["J"]))
    sd = ho.clone_state(net.state_dict())
    with torch.no_grad():
        out = ho.creat_model_s(sd, x, ho.Config(nStack=1, nOutChannels=16))
        loss = torch.nn.functional.mse_loss(out[0], tgt)
    np.testing.assert_array_equal(out[0].numpy(), g["out0"])
    assert loss.item() == float(g["loss"])


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("script", ["try_with_torch", "only_one_hourgless", "try_with_torch_100"])
def test_oracle_bit_exact_vs_reference_classes(script):
    ref = refload.load(script)
    old = (ref.nStack, ref.nOutChannels)
    try:
        ref.nStack, ref.nOutChannels = 2, 16
        torch.manual_seed(7)
        net = ref.creatModel()
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        x, tgt = model_inputs(8, 2, 128, 16)
        out = net(x)
        loss = sum(torch.nn.MSELoss()(o, tgt) for o in out)
        loss.backward()
        sd = ho.clone_state(sd0, requires_grad=True)
        out2 = ho.creat_model_s(sd, x, ho.Config(nStack=2, nOutChannels=16))
        tot, _ = ho.mse_losses(out2, tgt)
        tot.backward()
        assert all(torch.equal(a, b) for a, b in zip(out, out2))
        assert loss.item() == tot.item()
        for k, p in net.named_parameters():
            gr = sd[k].grad
            assert (p.grad is None) == (gr is None), k
            if gr is not None:
                assert torch.equal(p.grad, gr), k
        for k, v in net.state_dict().items():
            assert torch.equal(v, sd[k].detach()), k
    finally:
        ref.nStack, ref.nOutChannels = old


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_dropin_state_dict_layout_matches_reference():
    """Same keys, shapes and seeded values as the reference, so checkpoints interchange."""
    ref = refload.load("try_with_torch")
    torch.manual_seed(3)
    a = twt.creatModel().state_dict()
    torch.manual_seed(3)
    b = ref.creatModel().state_dict()
    assert list(a.keys()) == list(b.keys()) and len(a) == 199
    assert all(torch.equal(a[k], b[k]) for k in a)
    # module-global configuration is read at call time, like the reference (try_with_torch.py:224,285)
    twt.nStack = 8
    assert twt.creatModel()._config_key() == (8, 2)


@pytest.mark.skipif(not refload.available(), reason="reference tree not present (GPU box)")
def test_custom_losses_restate_reference():
    """oracle/losses_torch.py == the Costomer_* loss classes of train.py:343-408 (values and gradients, bit for bit)."""
    import warnings

    from oracle import losses_torch as lt
    tr = refload.load("train")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 17, 16, 16, generator=g)
    y = torch.randint(0, 17, (3, 16, 16), generator=g)
    t = torch.rand(3, 17, 16, 16, generator=g)
    mask = (torch.rand(3, 16, 16, generator=g) < 0.4)
    cases = [(tr.Costomer_CrossEntropyLoss(), lt.bootstrapped_cross_entropy, (y, 0.3)),
             (tr.Costomer_CrossEntropyLoss(), lt.bootstrapped_cross_entropy, (y, 0.01)),
             (tr.Costomer_CrossEntropyLoss_with_mask(), lt.masked_cross_entropy, (y, mask)),
             (tr.Costomer_MSELoss_with_mask(), lt.masked_mse, (t, mask)),
             (tr.Costomer_MSELoss(), lt.bootstrapped_mse, (t, 0.5)),
             (tr.Costomer_MSELoss(), lt.bootstrapped_mse, (t, 0.1))]
    for ref_mod, fn, args in cases:
        a = x.clone().requires_grad_()
        b = x.clone().requires_grad_()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            la = ref_mod.forward(a, *args)
        lb = fn(b, *args)
        la.backward()
        lb.backward()
        assert torch.equal(la, lb) and torch.equal(a.grad, b.grad), fn.__name__


def test_oracle_reproduces_warm_8stack_golden():
    """Headline configuration on WARM weights (tests/golden/warm_s_8stack.npz: 80 Adam steps of the real reference):
    the oracle port's train-mode forward and per-stack losses equal what the reference computed from those weights."""
    g = np.load(os.path.join(GOLDEN, "warm_s_8stack.npz"))

    def unbits(a):
        return torch.from_numpy(a.astype(np.int32)).to(torch.int32).bitwise_left_shift(16).view(torch.float32)

    sd = {}
    for k in (str(k) for k in g["keys"]):
        a = g["w:" + k]
        sd[k] = unbits(a) if a.dtype == np.uint16 else torch.from_numpy(a)
    twt.nStack, twt.nOutChannels = 8, 16
    assert list(twt.creatModel().state_dict().keys()) == list(sd.keys())
    with torch.no_grad():
        out = ho.creat_model_s(ho.clone_state(sd), unbits(g["x_bits"]), ho.Config(nStack=8, nOutChannels=16))
    tgt = torch.from_numpy(g["target"])
    for k in range(8):
        d = digest(out[k])
        # same torch ops on the same host: identical up to the thread count's reduction order
        np.testing.assert_allclose(d[2:], g["out_digest"][k][2:], rtol=2e-4, atol=1e-6)
        assert abs(torch.nn.functional.mse_loss(out[k], tgt).item() - g["losses"][k]) <= 1e-4 * g["losses"][k]
    np.testing.assert_allclose(out[7].numpy(), g["out7"], rtol=0, atol=2e-4 * np.abs(g["out7"]).max())
