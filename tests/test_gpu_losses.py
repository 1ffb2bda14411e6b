"""GPU parity of the bootstrapped / masked losses of train.py:343-408 (mirror classes in ...b200.train) and of the
radix-select top-k mask kernel, against oracle/losses_torch.py (pinned to the reference classes on the CPU) and
torch.topk.  Values rtol 1e-5 (north_star fp32); gradients rtol 1e-5 with an absolute term at 1e-5 of their scale."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import progressive_process_for_human_pose_estimation_b200 as hg  # noqa: E402
import progressive_process_for_human_pose_estimation_b200.train as tr  # noqa: E402
from oracle import losses_torch as lt  # noqa: E402
from progressive_process_for_human_pose_estimation_b200 import _lib as L  # noqa: E402


def close(got, want):
    return torch.allclose(got.cpu(), want, rtol=1e-5, atol=1e-5 * want.abs().max().item())


@pytest.mark.parametrize("n,k", [(4096, 409), (4096, 4096), (4096, 1), (69632, 1024), (1000, 250), (37, 5)])
def test_topk_mask_selects_exactly_the_k_largest(n, k):
    g = torch.Generator().manual_seed(n + k)
    v = torch.randn(5, n, generator=g)
    v[1] = v[1].abs()
    v[2] = torch.randint(0, 7, (n,), generator=g).float()          # massive ties at the threshold
    v[3] = 0.0                                                     # constant row: the first k indices win
    v[4, : n // 2] = -v[4, : n // 2].abs()
    vd = v.cuda()
    mask = torch.full_like(vd, -1.0)
    kth = torch.empty(5, device="cuda")
    L.call("hg_topk_mask", L.ptr(vd), 5, n, k, L.ptr(mask), L.ptr(kth), L.stream_ptr())
    m = mask.cpu()
    assert set(m.unique().tolist()) <= {0.0, 1.0}
    assert (m.sum(1) == k).all()
    top, _ = torch.topk(v, k, dim=1)
    assert torch.equal(kth.cpu(), top[:, -1])
    assert torch.equal((v * m).sum(1, dtype=torch.float64), top.sum(1, dtype=torch.float64)) or \
        torch.allclose((v * m).sum(1, dtype=torch.float64), top.sum(1, dtype=torch.float64), rtol=1e-12)
    for r in range(5):                                            # ties: lowest indices
        thr = top[r, -1]
        eq = (v[r] == thr).nonzero().flatten()
        chosen = eq[m[r, eq] == 1]
        assert torch.equal(chosen, eq[: len(chosen)])
    assert (m[3, :k] == 1).all() and (m[3, k:] == 0).all()


def test_bootstrapped_and_masked_losses_match_reference_statements():
    g = torch.Generator().manual_seed(1)
    B = 4
    x = (2 * torch.randn(B, 17, 64, 64, generator=g))
    y = torch.randint(0, 17, (B, 64, 64), generator=g)
    t = torch.rand(B, 17, 64, 64, generator=g)
    mask = torch.rand(B, 64, 64, generator=g) < 0.3
    cases = [(tr.Costomer_CrossEntropyLoss(), lt.bootstrapped_cross_entropy, (y, 0.3), 5),
             (tr.Costomer_CrossEntropyLoss(), lt.bootstrapped_cross_entropy, (y, 0.02), 5),   # clamps to 0.1
             (tr.Costomer_CrossEntropyLoss(), lt.bootstrapped_cross_entropy, (y, 1.0), 5),
             (tr.Costomer_CrossEntropyLoss_with_mask(), lt.masked_cross_entropy, (y, mask), 2),
             (tr.Costomer_MSELoss_with_mask(), lt.masked_mse, (t, mask), 1),
             (tr.Costomer_MSELoss(), lt.bootstrapped_mse, (t, 0.5), 3),
             (tr.Costomer_MSELoss(), lt.bootstrapped_mse, (t, 0.1), 3)]
    for mod, fn, args, launches in cases:
        a = x.clone().requires_grad_()
        ref = fn(a, *args)
        (3.0 * ref).backward()
        b = x.clone().cuda().requires_grad_()
        n0 = L.launch_count()
        got = mod.forward(b, *[v.cuda() if torch.is_tensor(v) else v for v in args])
        assert L.launch_count() - n0 == launches, (fn.__name__, L.launch_count() - n0)
        (3.0 * got).backward()
        assert got.shape == () and torch.allclose(got.cpu(), ref.detach(), rtol=1e-5), (fn.__name__, got.item(), ref.item())
        assert close(b.grad, a.grad), fn.__name__
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tr.Costomer_CrossEntropyLoss().forward(x, y, 0.5)


def test_bootstrapped_loss_trains_the_progressive_model():
    """train.py's objective on its own model: three bootstrapped cross-entropy heads (train.py:800-804,850-870) drive one
    Adam step of the drop-in creatModel; finite loss, gradients everywhere the reference has them."""
    torch.manual_seed(0)
    net = tr.creatModel().cuda()
    opt = hg.Adam(net.parameters(), lr=1e-4)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(2)).cuda()
    ys = [torch.randint(0, c, (2, 64, 64), generator=torch.Generator().manual_seed(3 + c)).cuda() for c in (2, 16, 17)]
    crit = tr.Costomer_CrossEntropyLoss()
    losses = []
    for _ in range(3):
        out = net(x)
        assert [tuple(o.shape) for o in out] == [(2, 2, 64, 64), (2, 16, 64, 64), (2, 17, 64, 64)]
        loss = sum(crit.forward(o, y, 0.25) for o, y in zip(out, ys))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
