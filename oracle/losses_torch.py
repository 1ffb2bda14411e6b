"""TEST INFRASTRUCTURE ONLY -- torch (CPU) restatement of the bootstrapped / masked losses of the reference's train.py,
statement for statement (the reference classes themselves cannot travel to the GPU box).  Pinned to the reference
classes by tests/test_oracle_model.py::test_custom_losses_restate_reference where /root/reference is present."""
import torch
import torch.nn.functional as F


def bootstrapped_cross_entropy(input, target, fraction):
    """Costomer_CrossEntropyLoss.forward, train.py:350-362."""
    if fraction < 0.1:
        fraction = 0.1
    loss = F.nll_loss(F.log_softmax(input, dim=1), target, reduction="none")
    k = input.shape[2] * input.shape[3] * fraction
    loss, _ = torch.topk(loss.view(input.shape[0], -1), int(k))
    return loss.mean()


def masked_cross_entropy(input, target, mask):
    """Costomer_CrossEntropyLoss_with_mask.forward, train.py:372-376."""
    loss = F.nll_loss(F.log_softmax(input, dim=1), target, reduction="none")
    loss = torch.mul(loss, mask.float()).view([loss.shape[0], -1])
    return loss.mean()


def masked_mse(input, target, mask):
    """Costomer_MSELoss_with_mask.forward, train.py:386-391."""
    loss = F.mse_loss(input, target, reduction="none")
    loss = torch.mul(loss, mask.float().view([mask.shape[0], 1, mask.shape[1], mask.shape[2]])).view([loss.shape[0], -1])
    return loss.mean()


def bootstrapped_mse(input, target, fraction):
    """Costomer_MSELoss.forward, train.py:401-408."""
    if fraction < 0.25:
        fraction = 0.25
    loss = F.mse_loss(input, target, reduction="none")
    k = input.shape[2] * input.shape[3] * fraction
    loss, _ = torch.topk(loss.view(input.shape[0], -1), int(k))
    return loss.mean()
