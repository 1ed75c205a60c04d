"""Host logic of the caller-side mirrors: constructor contracts, error behaviour and the refusal to run
without CUDA.  No kernel is launched here; the parity tests proper are tests/test_gpu_{loss,optim,metrics,voxel}.py."""
import numpy as np
import pytest
import torch


def test_charbonnier_constructor_contract():
    from refid_b200.losses import CharbonnierLoss
    c = CharbonnierLoss()
    assert (c.loss_weight, c.reduction, c.eps) == (1.0, "mean", 1e-12)  # losses.py:155
    with pytest.raises(ValueError, match="Unsupported reduction mode"):
        CharbonnierLoss(reduction="avg")
    with pytest.raises(RuntimeError, match="no CPU path"):
        c(torch.zeros(4, requires_grad=True), torch.zeros(4))
    with pytest.raises(NotImplementedError):
        c(torch.zeros(4), torch.zeros(4), weight=torch.ones(4))


def test_clip_adam_constructor_and_state_layout():
    from refid_b200 import optim
    p = torch.nn.Parameter(torch.zeros(3))
    o = optim.ClipAdamW([p], lr=2e-4, betas=(0.9, 0.99), weight_decay=1e-4)
    assert isinstance(o, torch.optim.Optimizer)
    g = o.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"]) == (2e-4, (0.9, 0.99), 1e-8, 1e-4)
    assert optim.ClipAdam([p]).param_groups[0]["weight_decay"] == 0  # torch.optim.Adam's default
    with pytest.raises(ValueError):
        optim.ClipAdamW([p], lr=-1.0)
    with pytest.raises(NotImplementedError):
        optim.ClipAdamW([p], amsgrad=True)
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        o.step()
    sd = o.state_dict()
    assert sd["param_groups"][0]["lr"] == 2e-4


def test_metrics_and_event_util_need_cuda():
    from refid_b200 import event_util, metrics
    with pytest.raises(RuntimeError, match="no CPU path"):
        metrics.psnr_frames(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8))
    with pytest.raises(TypeError):
        metrics.tensor2img(np.zeros((3, 8, 8)))
    with pytest.raises(RuntimeError, match="no CPU path"):
        event_util.events_to_voxel_grid(torch.zeros(4, 4), 2, 8, 8)


def test_event_oracle_properties():
    """Size-independent properties of the rasterisation the GPU tests rely on: the grid sums to the sum of polarities over
    the events whose upper bin is valid, and the two-bin weights of an event add up to its polarity."""
    from oracle import event_oracle as E
    ev = E.synthetic_events(2000, 32, 24, seed=9)
    v = E.events_to_voxel_grid(ev, 2, 32, 24)
    pol = np.where(ev[:, 3] == 0, -1.0, 1.0)
    # every event but those at the very last stamp has ti = 0: both weights land in the grid and sum to the polarity
    last = ev[:, 0] == ev[-1, 0]
    assert abs(v.sum() - (pol[~last].sum() + pol[last].sum())) < 1e-3
    wins = E.sliding_two_bin_voxels([ev[:700], ev[700:1400], ev[1400:]], 32, 24)
    assert len(wins) == 2 and wins[0].shape == (24, 32, 2)
