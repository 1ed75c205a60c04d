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


def test_optimizer_state_interchanges_with_the_reference_group_layout():
    """The reference builds AdamW over TWO groups, the second (`optim_params_lowlr`, lr x 0.1) empty for this network
    (twoImage_event_recurrent_model.py:67-91).  A state_dict produced by that layout loads into the mirror's optimizer and
    the mirror's state_dict loads into a reference-shaped torch optimizer (resuming from a reference .state)."""
    from refid_b200 import plugin, recurrent_model
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    opt = plugin.load_options(os.path.join(root, "options/train/GoPro_blurry_11p1_b200.yml"))
    m = recurrent_model.TwoImageEventRecurrentRestorationModel(opt, device="cpu")
    ours = m.optimizer_g
    assert [len(g["params"]) for g in ours.param_groups] == [183, 0]
    assert ours.param_groups[1]["lr"] == pytest.approx(ours.param_groups[0]["lr"] * 0.1)
    params = [p for p in m.net_g.parameters()]
    og = dict(opt["train"]["optim_g"])
    og.pop("type")
    ref = torch.optim.AdamW([{"params": params}, {"params": [], "lr": og["lr"] * 0.1}], **og)  # the reference's call
    for p in params[:5]:  # a few populated states, as after training steps
        ref.state[p] = {"step": torch.tensor(7.0), "exp_avg": torch.full_like(p, 0.5), "exp_avg_sq": torch.full_like(p, 0.25)}
    ours.load_state_dict(ref.state_dict())
    st = ours.state[params[0]]
    assert float(st["step"]) == 7.0 and torch.equal(st["exp_avg"], torch.full_like(params[0], 0.5))
    ref2 = torch.optim.AdamW([{"params": params}, {"params": [], "lr": og["lr"] * 0.1}], **og)
    ref2.load_state_dict(ours.state_dict())  # and back
    assert float(ref2.state[params[4]]["step"]) == 7.0
    assert [len(g["params"]) for g in ours.state_dict()["param_groups"]] == [183, 0]


def test_two_optimizers_of_different_sizes_share_the_library_prototype():
    """refid_optim_step's ctypes prototype is pointer-typed, not sized by the first optimizer's tensor count."""
    import inspect
    from refid_b200 import optim
    src = inspect.getsource(optim._ClipAdamBase.step)
    assert "ctypes.c_long * n," not in src and "ctypes.cast" in src


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
