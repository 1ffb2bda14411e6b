"""GPU parity tests of the drop-in modules (whole plans through the C ABI) against the CPU oracle
(oracle/hourglass_torch.py, itself pinned bit-for-bit to the reference: tests/test_oracle_model.py).

Protocol: the seeded-init / train-mode-BN network amplifies rounding ~x3.5 per stack, so tight
tolerances are asserted (i) per block with teacher forcing, (ii) end to end in eval() mode and on the fp32 path;
the bf16 end-to-end train-mode run is checked at the level the reference itself reproduces (loss, statistics).
Tolerances: fp32 path rtol 1e-5 per op (north_star) -> 1e-4 after a few dozen chained ops; bf16 rtol 2e-2.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

import progressive_process_for_human_pose_estimation_b200 as hg  # noqa: E402
import progressive_process_for_human_pose_estimation_b200.try_with_torch as m  # noqa: E402
from oracle import hourglass_torch as ho  # noqa: E402


@pytest.fixture(autouse=True)
def _defaults():
    m.nStack, m.nModules, m.nFeats, m.nOutChannels = 4, 2, 256, 17
    yield
    hg.set_compute_dtype(torch.bfloat16)
    m.nStack, m.nModules, m.nFeats, m.nOutChannels = 4, 2, 256, 17


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def _sd_prefix(sd, prefix):
    return {prefix + "." + k: v for k, v in sd.items()}


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("numIn,numOut,hw", [(256, 256, 16), (64, 128, 32), (128, 128, 8)])
def test_residual_block_teacher_forced(dtype, tol, numIn, numOut, hw):
    hg.set_compute_dtype(dtype)
    torch.manual_seed(0)
    blk = m.ResidualBlock(numIn, numOut)
    # non-trivial BN affine parameters
    for bn in (blk.bn1, blk.bn2, blk.bn3):
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.2)
    sd0 = copy.deepcopy(blk.state_dict())
    x = torch.randn(8, numIn, hw, hw)
    gout = torch.randn(8, numOut, hw, hw)
    if dtype == torch.bfloat16:  # teacher forcing: both sides see the same representable inputs
        x, gout = x.bfloat16().float(), gout.bfloat16().float()
    sd = ho.clone_state(_sd_prefix(sd0, "b"), requires_grad=True)
    xo = x.clone().requires_grad_(True)
    yo = ho.residual_block(sd, "b", xo, numIn, numOut, True)
    yo.backward(gout)
    # Gradient yardstick for bf16: ReLU masks flip where a bf16-rounded pre-activation sits at ~0, so the
    # reference ITSELF under torch.autocast(bfloat16) is several % away from its fp32 gradients.  The bf16 kernels
    # must be as close to the fp32 oracle as the reference's own bf16 execution (x2), never worse than that.
    gtol = {"__dx__": 2 * tol}
    if dtype == torch.bfloat16:
        sda = ho.clone_state(_sd_prefix(sd0, "b"), requires_grad=True)
        xa = x.clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            ya = ho.residual_block(sda, "b", xa, numIn, numOut, True)
        ya.float().backward(gout)
        gtol["__dx__"] = max(2 * tol, 2 * rel(xa.grad, xo.grad))
        for k, v in sda.items():
            if v.grad is not None and sd[k].grad is not None:
                gtol[k] = max(4 * tol, 2 * rel(v.grad.float(), sd[k].grad))
    blk = blk.cuda()
    xc = x.cuda().requires_grad_(True)
    for it in range(3):  # iteration 0 eager, 1 eager bwd / graph fwd, 2 graphs
        blk.load_state_dict(sd0)
        blk.zero_grad(set_to_none=True)
        xc.grad = None
        y = blk(xc)
        y.backward(gout.cuda())
        assert maxrel(y.detach().cpu(), yo.detach()) <= tol, f"forward (iter {it})"
        assert rel(xc.grad.cpu(), xo.grad) <= gtol["__dx__"], f"input gradient (iter {it})"
        gmax = max(p.grad.abs().max().item() for n_, p in blk.named_parameters() if p.grad is not None)
        for name, p in blk.named_parameters():
            go = sd["b." + name].grad
            if p.grad is None:
                assert go is None
                continue
            if "conv1.bias" in name or "conv2.bias" in name:
                # a bias feeding a BatchNorm has an analytically zero gradient (oracle: rounding noise ~1e-7);
                # only an absolute bound is meaningful
                assert p.grad.abs().max().item() <= (5e-2 if dtype == torch.bfloat16 else 1e-4) * gmax, name
                continue
            assert rel(p.grad.cpu(), go) <= gtol.get("b." + name, 4 * tol), f"grad of {name} (iter {it})"
        for k, v in blk.state_dict().items():
            if "running" in k:
                assert maxrel(v.cpu(), sd["b." + k].detach()) <= max(tol, 1e-5), k
            if "num_batches" in k:
                assert int(v) == int(sd["b." + k])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 4e-2)])
def test_hourglass_module_vs_oracle(dtype, tol):
    """hourglass(2, 256) on a 16x16 map: pooling, two levels of shared blocks, bilinear up-sampling + add."""
    hg.set_compute_dtype(dtype)
    torch.manual_seed(1)
    mod = m.hourglass(2, 256)
    sd0 = copy.deepcopy(mod.state_dict())
    x = torch.randn(16, 256, 16, 16).bfloat16().float()
    gout = torch.randn(16, 256, 16, 16).bfloat16().float()
    sd = ho.clone_state(_sd_prefix(sd0, "hourglass1"), requires_grad=True)
    xo = x.clone().requires_grad_(True)
    yo = ho.hourglass(sd, "hourglass1", xo, 2, ho.Config())
    yo.backward(gout)
    gscale = 1.0
    if dtype == torch.bfloat16:  # yardstick: the reference's own autocast-bf16 divergence (see the block test)
        sda = ho.clone_state(_sd_prefix(sd0, "hourglass1"), requires_grad=True)
        xa = x.clone().requires_grad_(True)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            ya = ho.hourglass(sda, "hourglass1", xa, 2, ho.Config())
        ya.float().backward(gout)
        tol = max(tol, 1.5 * rel(ya.detach().float(), yo.detach()))
        gscale = max(1.0, 2 * rel(xa.grad, xo.grad) / (5 * tol))
    mod = mod.cuda()
    xc = x.cuda().requires_grad_(True)
    y = mod(xc)
    y.backward(gout.cuda())
    assert rel(y.detach().cpu(), yo.detach()) <= tol
    # 14 chained train-mode blocks: the order of the fp32 statistics atomics changes run to run and the backward pass
    # amplifies that last-bit noise (observed 0.5e-3 .. 2e-3 on the fp32 path); per-block parity is the tight test
    assert rel(xc.grad.cpu(), xo.grad) <= 30 * tol * gscale
    worst = 0.0
    for name, p in mod.named_parameters():
        go = sd["hourglass1." + name].grad
        if p.grad is None:
            assert go is None or go.abs().max() == 0
            continue
        if "conv1.bias" in name or "conv2.bias" in name:
            continue
        worst = max(worst, rel(p.grad.cpu(), go))
    assert worst <= 30 * tol * gscale, worst
    nb = {k: int(v) for k, v in mod.state_dict().items() if "num_batches" in k}
    assert all(nb[k] == int(sd["hourglass1." + k]) for k in nb)


def _model_and_oracle(nStack, J, B, training, seed=0):
    m.nStack, m.nOutChannels = nStack, J
    torch.manual_seed(seed)
    net = m.creatModel()
    if not training:  # give eval mode non-trivial running statistics
        g = torch.Generator().manual_seed(5)
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
    sd0 = copy.deepcopy(net.state_dict())
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, 3, 256, 256, generator=g)
    tgt = torch.rand(B, J, 64, 64, generator=g)
    cfg = ho.Config(nStack=nStack, nOutChannels=J, training=training)
    return net, sd0, x, tgt, cfg


def test_model_fp32_train_step_vs_oracle():
    """BASELINE config 1 shape (1 stack, 16 heatmaps) on the fp32 path: outputs, loss, gradients, BN buffers."""
    hg.set_compute_dtype(torch.float32)
    net, sd0, x, tgt, cfg = _model_and_oracle(1, 16, 2, True)
    sd = ho.clone_state(sd0, requires_grad=True)
    oo = ho.creat_model_s(sd, x, cfg)
    tot, _ = ho.mse_losses(oo, tgt)
    tot.backward()
    net = net.cuda()
    out = net(x.cuda())
    loss = sum(torch.nn.MSELoss()(o, tgt.cuda()) for o in out)
    loss.backward()
    assert isinstance(out, list) and len(out) == 1 and out[0].shape == (2, 16, 64, 64)
    assert rel(out[0].cpu(), oo[0].detach()) <= 2e-4
    assert abs(loss.item() - tot.item()) <= 1e-5 * abs(tot.item())
    unused = 0
    for name, p in net.named_parameters():
        go = sd[name].grad
        if p.grad is None:
            unused += 1
            assert go is None or go.abs().max() == 0, name
            continue
        if ".conv1.bias" in name or ".conv2.bias" in name or name == "lin.conv.bias":
            continue  # feeds a BatchNorm: analytically zero gradient, rounding noise on both sides
        if go.norm() > 1e-6:  # chaotic conditioning at random init (Q13): loose relative, see block tests for tight
            assert rel(p.grad.cpu(), go) <= 0.05, (name, rel(p.grad.cpu(), go))
    # grad-less tensors: conv4 of the 6 identity blocks (quirk Q3) + the re-injection convs after the last stack (Q5)
    assert unused == 16 == sum(1 for k, v in sd.items() if v.requires_grad and v.grad is None)
    for k, v in net.state_dict().items():
        if "running" in k:
            assert maxrel(v.cpu(), sd[k].detach()) <= 1e-3, k
        if "num_batches" in k:
            assert int(v) == int(sd[k])


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_model_eval_mode_4stack_vs_oracle(dtype, tol):
    """BASELINE config 5 network (4 stacks, 17 COCO heatmaps) in eval(): well conditioned, so bf16 holds 2-3e-2."""
    hg.set_compute_dtype(dtype)
    net, sd0, x, tgt, cfg = _model_and_oracle(4, 17, 2, False)
    sd = ho.clone_state(sd0)
    with torch.no_grad():
        oo = ho.creat_model_s(sd, x, cfg)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(x.cuda())
        out2 = net(x.cuda())  # CUDA-graph replay
    assert len(out) == 4
    plan = next(iter(net._plans.values()))
    if dtype == torch.bfloat16:  # inference: BN2/BN3 of every residual block and lin.bn run in their producer conv
        assert plan.n_bn_out > 0 and any(c.name == "hg_conv_fprop_bnout" for c in plan.fwd_calls)
    for k in range(4):
        assert rel(out[k].cpu(), oo[k]) <= tol * (1 + k), (k, rel(out[k].cpu(), oo[k]))
        assert torch.equal(out[k], out2[k])
    for k, v in net.state_dict().items():  # eval never touches the buffers
        assert torch.equal(v.cpu(), sd0[k])


def test_model_bf16_train_statistics_and_graph_replay():
    """bf16 train-mode step: loss close to the oracle's, BN bookkeeping exact, graph replay deterministic."""
    hg.set_compute_dtype(torch.bfloat16)
    net, sd0, x, tgt, cfg = _model_and_oracle(2, 16, 4, True)
    sd = ho.clone_state(sd0, requires_grad=True)
    oo = ho.creat_model_s(sd, x, cfg)
    tot, per = ho.mse_losses(oo, tgt)
    net = net.cuda()
    xc, tc = x.cuda(), tgt.cuda()
    losses, grads = [], []
    for it in range(3):
        net.load_state_dict(sd0)
        net.zero_grad(set_to_none=True)
        out = net(xc)
        loss = sum(torch.nn.MSELoss()(o, tc) for o in out)
        loss.backward()
        losses.append(loss.item())
        grads.append(net.conv2.weight.grad.clone())
    assert abs(losses[0] - tot.item()) <= 2e-2 * tot.item()
    # run-to-run: fp32 atomics reorder the BN statistics sums; the chaotic net (Q13) amplifies that last-bit noise
    assert abs(losses[1] - losses[0]) <= 1e-2 * losses[0] and abs(losses[2] - losses[0]) <= 1e-2 * losses[0]
    assert torch.isfinite(grads[2]).all() and grads[2].abs().max() > 0
    nbt = {k: int(v) for k, v in net.state_dict().items() if "num_batches" in k}
    assert all(nbt[k] == int(sd[k]) for k in nbt)
    f, b = net.launches_per_step()
    assert f > 0 and b > 0


def test_state_dict_roundtrip_and_frozen_param():
    m.nStack, m.nOutChannels = 1, 16
    torch.manual_seed(0)
    net = m.creatModel().cuda()
    sd = net.state_dict()
    assert len(sd) == 199
    net2 = m.creatModel().cuda()
    net2.load_state_dict(sd)
    net2.conv3.weight.requires_grad_(False)
    x = torch.randn(1, 3, 256, 256, device="cuda")
    out = net2(x)
    out[0].square().mean().backward()
    assert net2.conv3.weight.grad is None and net2.conv2.weight.grad is not None


def test_cpu_input_fails_loudly():
    m.nStack = 1
    net = m.creatModel()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.randn(1, 3, 256, 256))


def test_full_size_train_steps_properties():
    """BASELINE configs[1] at full size (8 stacks, 16 joints, B=32, bf16): properties that do not need an oracle --
    finite decreasing loss over Adam steps, exact BatchNorm bookkeeping (a level's shared BN modules are updated
    6-8 x nStack times per forward, quirk Q1), CUDA-graph replays, gradients only where the reference has them."""
    hg.set_compute_dtype(torch.bfloat16)
    m.nStack, m.nOutChannels = 8, 16
    torch.manual_seed(0)
    net = m.creatModel().cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(32, 3, 256, 256, generator=g).cuda()
    import numpy as np
    r = np.random.RandomState(4)
    kp = np.zeros([32, 1, 16, 3])
    kp[..., 0], kp[..., 1], kp[..., 2] = r.randint(0, 640, [32, 1, 16]), r.randint(0, 480, [32, 1, 16]), 2
    y = hg.gaussian_heatmaps(kp, np.tile(np.array([[640.0, 480.0]]), (32, 1)))
    losses = []
    steps = 10  # the first Adam steps at lr 1e-3 overshoot (the loss rises before it falls); atomics make runs differ
    for it in range(steps):
        out = net(x)
        assert len(out) == 8 and all(o.shape == (32, 16, 64, 64) for o in out)
        loss = sum(torch.nn.MSELoss()(o, y) for o in out)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < 0.7 * losses[0], losses
    sd = net.state_dict()
    assert int(sd["hourglass1.residual_block.bn1.num_batches_tracked"]) == 6 * 8 * steps  # 6 calls x 8 stacks x steps
    assert int(sd["hourglass1.hourglass1.hourglass1.hourglass1.residual_block.bn1.num_batches_tracked"]) == 8 * 8 * steps
    assert int(sd["residual4.bn2.num_batches_tracked"]) == 2 * 8 * steps
    assert int(sd["residual1.bn1.num_batches_tracked"]) == steps
    none = [n for n, p in net.named_parameters() if p.grad is None]
    assert len(none) == 12 and all(".conv4." in n for n in none)   # identity blocks' unused projections (quirk Q3)
    # decode of the (identical) targets is exact at full size
    yx, mx = hg.decode_argmax(y)
    assert torch.equal(yx[..., 0].long() * 64 + yx[..., 1].long(), y.flatten(2).argmax(-1))


def test_half_model_like_the_reference_test_mode():
    """The reference's test branches run `model.half()` WITHOUT .eval() (try_with_torch.py:374-377, quirk Q10): fp16
    parameters and BN buffers, batch statistics at batch 1.  The drop-in takes the same calls; results against the
    fp32 oracle at bf16-path tolerance, BN bookkeeping carried by the (half) buffers."""
    hg.set_compute_dtype(torch.bfloat16)
    net, sd0, x, tgt, cfg = _model_and_oracle(2, 17, 1, True)
    net = net.half().cuda()
    sdh = {k: (v.half().float() if v.is_floating_point() else v.clone()) for k, v in sd0.items()}  # fp16-rounded weights
    sd = ho.clone_state(sdh)
    with torch.no_grad():
        oo = ho.creat_model_s(sd, x.half().float(), cfg)
        out = net(x.half().cuda())
        out_again = net(x.half().cuda())  # CUDA-graph replay
    assert all(o.dtype == torch.float16 for o in out)
    # train-mode BN at batch 1, random init: chaotic (Q13) -- a sanity bound only (0.29 .. 0.31 measured from run to
    # run: the BatchNorm sums arrive in a different order); the tight comparison of this model is the eval() part below
    assert rel(out[0].float().cpu(), oo[0]) <= 0.5
    assert torch.isfinite(out_again[1].float()).all()
    st = net.state_dict()
    assert st["residual1.bn1.running_mean"].dtype == torch.float16
    assert int(st["residual1.bn1.num_batches_tracked"]) == 2
    assert not torch.equal(st["residual1.bn1.running_mean"].float().cpu(), sd0["residual1.bn1.running_mean"])
    # eval() on the same half model: well conditioned
    net.load_state_dict({k: v for k, v in sdh.items()})
    net.eval()
    cfg.training = False
    with torch.no_grad():
        oe = ho.creat_model_s(ho.clone_state(sdh), x.half().float(), cfg)
        oute = net(x.half().cuda())
    for k in range(2):
        assert rel(oute[k].float().cpu(), oe[k]) <= 3e-2 * (1 + k), (k, rel(oute[k].float().cpu(), oe[k]))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_model_eval_mode_8stack_headline_vs_oracle(dtype, tol):
    """The HEADLINE network (BASELINE configs[1]: try_with_torch.creatModel, nStack=8, 16 heatmaps) in eval() at B=2
    against the oracle port, per stack: fp32 <= 1e-4 (1+k), bf16 <= 2e-2 (1+k).  Catches a wrong shared-weight call
    site or accumulation in stacks 5-8, which the 4-stack test cannot see."""
    hg.set_compute_dtype(dtype)
    net, sd0, x, tgt, cfg = _model_and_oracle(8, 16, 2, False)
    sd = ho.clone_state(sd0)
    with torch.no_grad():
        oo = ho.creat_model_s(sd, x, cfg)
    net = net.cuda().eval()
    with torch.no_grad():
        out = net(x.cuda())
        out2 = net(x.cuda())  # CUDA-graph replay
    assert len(out) == 8 and all(o.shape == (2, 16, 64, 64) for o in out)
    for k in range(8):
        assert rel(out[k].cpu(), oo[k]) <= tol * (1 + k), (k, rel(out[k].cpu(), oo[k]))
        assert torch.equal(out[k], out2[k])


def _load_warm():
    import os

    import numpy as np

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warm_s_8stack.npz"))
    sd = {}
    for k in (str(k) for k in g["keys"]):
        a = g["w:" + k]
        if a.dtype == np.uint16:  # bf16 bit patterns
            sd[k] = torch.from_numpy(a.astype(np.int32)).to(torch.int32).bitwise_left_shift(16).view(torch.float32)
        else:
            sd[k] = torch.from_numpy(a)
    x = torch.from_numpy(g["x_bits"].astype(np.int32)).to(torch.int32).bitwise_left_shift(16).view(torch.float32)
    return g, sd, x, torch.from_numpy(g["target"])


def _digest(t):
    f = t.detach().double().reshape(-1).cpu()
    idx = torch.linspace(0, f.numel() - 1, 24).long()
    return torch.cat([torch.stack([f.sum(), f.abs().sum()]), f[idx]]).numpy()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_warm_weight_8stack_train_step_vs_reference_golden(dtype):
    """SURVEY Q13 protocol (ii) on the HEADLINE configuration: weights after 80 fp32 Adam steps of the REAL reference
    (oracle/make_golden_warm.py; tests/golden/warm_s_8stack.npz), one TRAIN-mode step (batch statistics, 8 MSE terms,
    backward) at B=2 against what the reference computed from the same weights: per-stack heatmaps, per-stack losses,
    every parameter gradient, BatchNorm buffers.  fp32 path: 15 x the reference's own fp32-vs-fp64 divergence (5e-6 at
    stack 1 .. 8e-5 at stack 8); bf16 path: 1.5 x the reference's own autocast(bfloat16)-vs-fp32 divergence stored in
    the fixture (the trained heatmaps are ~0, so bf16 noise is 10-36 % of their norm for the reference itself)."""
    import numpy as np

    g, sd, x, tgt = _load_warm()
    hg.set_compute_dtype(dtype)
    m.nStack, m.nOutChannels = 8, 16
    net = m.creatModel()
    net.load_state_dict(sd)
    net = net.cuda().train()
    out = net(x.cuda())
    losses = [torch.nn.MSELoss()(o, tgt.cuda()) for o in out]
    sum(losses).backward()
    y64, ybf = g["yardstick_fp64"], g["yardstick_bf16"]
    fp32 = dtype == torch.float32
    for k in range(8):
        tol = max(1e-4, 15 * float(y64[k])) if fp32 else max(2e-2, 1.5 * float(ybf[k]))
        d = _digest(out[k])
        err = np.linalg.norm(d[2:] - g["out_digest"][k][2:]) / np.linalg.norm(g["out_digest"][k][2:])
        assert err <= 3 * tol, (k, err, tol)                       # 24 samples of every stack
        ltol = tol if fp32 else max(2e-2, float(ybf[k]) ** 2 * 4)  # an MSE moves with the square of the output error
        assert abs(losses[k].item() - g["losses"][k]) <= ltol * g["losses"][k], (k, losses[k].item(), g["losses"][k])
    for k, name in ((0, "out0"), (3, "out3"), (7, "out7")):          # full tensors of three stacks
        tol = max(1e-4, 15 * float(y64[k])) if fp32 else max(2e-2, 1.5 * float(ybf[k]))
        assert rel(out[k].detach().cpu(), torch.from_numpy(g[name])) <= tol, (name, tol)
    names = [str(n) for n in g["param_names"]]
    params = dict(net.named_parameters())
    assert list(params) == names
    gnorm = g["grad_norm"]
    small = 1e-4 * np.median(gnorm[gnorm > 0])
    gtol = 2e-2 if fp32 else 0.5
    checked, worst = 0, 0.0
    for i, n in enumerate(names):
        p = params[n]
        if g["grad_is_none"][i]:
            assert p.grad is None or p.grad.abs().max().item() == 0, n
            continue
        assert p.grad is not None and torch.isfinite(p.grad).all(), n
        if gnorm[i] < small:
            continue   # analytically-zero gradients (conv biases that feed a BatchNorm): rounding noise on both sides
        e = abs(p.grad.double().norm().item() - gnorm[i]) / gnorm[i]
        worst = max(worst, e)
        # yardstick per tensor: the reference's own fp32-vs-fp64 gradient divergence on these weights (median 1.8 %:
        # eight weight-shared stacks amplify rounding in the backward pass too; the stem's gradient, a cancelling sum
        # over every block of every stack, is the noisiest)
        assert e <= max(gtol, 4 * float(g["grad_noise_fp64"][i])), (n, e, float(g["grad_noise_fp64"][i]))
        checked += 1
    assert checked > 60, checked   # 199 tensors - 28 grad-less - the analytically-zero bias gradients
    if fp32:
        for n in (str(s) for s in g["small_grad_names"]):
            i = names.index(n)
            if gnorm[i] < small:
                continue
            tol_n = max(2e-2, 4 * float(g["grad_noise_fp64"][i]))
            assert rel(params[n].grad.cpu(), torch.from_numpy(g["g:" + n])) <= tol_n, (n, tol_n)
    sdn = net.state_dict()
    for i, k in enumerate(str(k) for k in g["keys"]):
        if "num_batches_tracked" in k:
            assert _digest(sdn[k].float())[2] == g["after_digest"][i][2], k
        elif "running" in k and fp32:
            a, b = _digest(sdn[k].float())[2:], g["after_digest"][i][2:]
            assert np.abs(a - b).max() <= 2e-3 * np.abs(b).max() + 1e-6, k
