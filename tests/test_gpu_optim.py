"""hg.Adam (one-launch Adam, csrc/optim.cu) against torch.optim.Adam -- the optimizer of every reference training loop
(try_with_torch.py:317,342-344) -- run by the stock class on the CPU in fp32 on identical parameters and gradients."""
import pytest
import torch

import progressive_process_for_human_pose_estimation_b200 as hg


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 3, 7, 7), (64,), (128, 64, 1, 1), (17, 256, 1, 1), (17,), (128, 128, 3, 3), (5,), (33000,), (1,)]
    return [torch.randn(s, generator=g) for s in shapes], g


def test_state_dict_interoperates_with_torch_adam_cpu():
    ps, _ = _params(0)
    ours = hg.Adam([torch.nn.Parameter(p.clone()) for p in ps], lr=1e-4, eps=1e-4)
    stock = torch.optim.Adam([torch.nn.Parameter(p.clone()) for p in ps], lr=1e-4, eps=1e-4)
    stock.load_state_dict(ours.state_dict())
    ours.load_state_dict(stock.state_dict())
    assert ours.param_groups[0]["eps"] == 1e-4
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        hg.Adam([p]).step()
    with pytest.raises(RuntimeError, match="amsgrad"):
        hg.Adam([p], amsgrad=True)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(lr=1e-4), dict(lr=1e-3, eps=1e-4), dict(lr=2e-3, betas=(0.8, 0.99), weight_decay=1e-2)])
def test_adam_matches_torch_adam(kw):
    ps, g = _params(1)
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    dev = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    o_ref, o_dev = torch.optim.Adam(ref, **kw), hg.Adam(dev, **kw)
    n0 = hg._lib.launch_count()
    for it in range(12):
        for r, d in zip(ref, dev):
            gr = torch.randn(r.shape, generator=g) * (0.1 + it)
            r.grad, d.grad = gr.clone(), gr.cuda()
        o_ref.step()
        o_dev.step()
        if it == 5:   # checkpoint round trip through the stock optimizer's format (try_with_torch.py:324-328,361-367)
            sd = o_dev.state_dict()
            o_dev = hg.Adam(dev, **kw)
            o_dev.load_state_dict(sd)
    assert hg._lib.launch_count() - n0 == 12   # one launch per step for all 9 tensors
    for r, d in zip(ref, dev):
        assert torch.allclose(d.detach().cpu(), r.detach(), rtol=1e-5, atol=1e-7), (d.detach().cpu() - r.detach()).abs().max()
        sr, sd_ = o_ref.state[r], o_dev.state[d]
        assert float(sd_["step"]) == float(sr["step"]) == 12
        assert torch.allclose(sd_["exp_avg"].cpu(), sr["exp_avg"], rtol=1e-5, atol=1e-9)
        assert torch.allclose(sd_["exp_avg_sq"].cpu(), sr["exp_avg_sq"], rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
def test_adam_on_the_dropin_model_gradients():
    """hg.Adam driven by the drop-in model's own gradients (views of the plan's flat gradient arena) for 3 training
    steps; a stock torch.optim.Adam fed the SAME gradients on a twin parameter set must end at the same weights."""
    import progressive_process_for_human_pose_estimation_b200.try_with_torch as m
    m.nStack, m.nOutChannels = 1, 16
    hg.set_compute_dtype(torch.float32)
    try:
        x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(2)).cuda()
        y = torch.rand(2, 16, 64, 64, generator=torch.Generator().manual_seed(3)).cuda()
        torch.manual_seed(0)
        net = m.creatModel().cuda()
        opt = hg.Adam(net.parameters(), lr=1e-4)
        named = [(k, p) for k, p in net.named_parameters()]
        twins = [torch.nn.Parameter(p.detach().clone()) for _, p in named]
        stock = torch.optim.Adam(twins, lr=1e-4)
        losses = []
        for _ in range(3):
            loss = sum(torch.nn.MSELoss()(o, y) for o in net(x))
            opt.zero_grad()
            loss.backward()
            for (_, p), t in zip(named, twins):
                t.grad = None if p.grad is None else p.grad.detach().clone()
            opt.step()
            stock.step()
            losses.append(loss.item())
        assert losses[-1] < losses[0]
        for (k, p), t in zip(named, twins):
            assert torch.allclose(p.detach(), t.detach(), rtol=1e-5, atol=1e-7), k
    finally:
        hg.set_compute_dtype(torch.bfloat16)
        m.nStack, m.nOutChannels = 4, 17
