"""Charbonnier loss kernel against the oracle restatement of basicsr/models/losses/losses.py:28-30."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,reduction,weight", [((2, 3, 3, 32, 40), "mean", 1.0), ((1, 2, 3, 17, 19), "sum", 0.5),
                                                    ((8, 23, 3, 256, 256), "mean", 1.0)])
def test_charbonnier_matches_oracle(shape, reduction, weight):
    from oracle import refid_oracle as O
    from refid_b200.losses import CharbonnierLoss
    g = torch.Generator().manual_seed(5)
    pred = torch.rand(shape, generator=g)
    gt = torch.rand(shape, generator=g)
    gt.view(-1)[::7] = pred.view(-1)[::7]  # exact zeros of the residual: the gradient there is 0 / sqrt(eps) = 0
    p_ref = pred.clone().requires_grad_(True)
    ref = O.charbonnier(p_ref, gt)
    if reduction == "sum":
        ref = ref * pred.numel()
    ref = weight * ref
    ref.backward()
    p = pred.cuda().requires_grad_(True)
    cri = CharbonnierLoss(loss_weight=weight, reduction=reduction)
    loss = cri(p, gt.cuda())
    (2.0 * loss).backward()  # upstream gradient != 1 exercises the device-side scale
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    assert torch.allclose(p.grad.cpu(), 2.0 * p_ref.grad, rtol=1e-4, atol=1e-9)
    # bit-reproducible, and the common upstream gradient of exactly 1 leaves the gradient untouched
    p2 = pred.cuda().requires_grad_(True)
    loss2 = cri(p2, gt.cuda())
    loss2.backward()
    assert loss2.item() == loss.item()
    assert torch.allclose(p2.grad.cpu(), p_ref.grad, rtol=1e-4, atol=1e-9)


def test_charbonnier_interface_errors():
    from refid_b200.losses import CharbonnierLoss
    with pytest.raises(ValueError):
        CharbonnierLoss(reduction="median")
    with pytest.raises(NotImplementedError):
        CharbonnierLoss(reduction="none")(torch.zeros(4, device="cuda"), torch.zeros(4, device="cuda"))
    with pytest.raises(RuntimeError):
        CharbonnierLoss()(torch.zeros(4, requires_grad=True), torch.zeros(4))
