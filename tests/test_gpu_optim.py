"""Fused clip + AdamW / Adam step against torch.nn.utils.clip_grad_norm_ + torch.optim on the CPU
(the reference's own calls, twoImage_event_recurrent_model.py:304-307)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 32, 3, 3), (1,), (128, 257), (3,), (40000,), (1, 64, 1, 1), (7, 5)]


@pytest.mark.parametrize("kind,max_norm", [("AdamW", 0.01), ("AdamW", 0.0), ("Adam", 0.5)])
def test_clip_adam_matches_torch(kind, max_norm):
    from refid_b200 import optim
    g = torch.Generator().manual_seed(11)
    P = [torch.randn(s, generator=g) for s in SHAPES]
    hp = dict(lr=2e-4, betas=(0.9, 0.99), weight_decay=1e-4)
    ref = [p.clone().requires_grad_(True) for p in P]
    mine = [p.clone().cuda().requires_grad_(True) for p in P]
    o_ref = (torch.optim.AdamW if kind == "AdamW" else torch.optim.Adam)(ref, **hp)
    o_mine = (optim.ClipAdamW if kind == "AdamW" else optim.ClipAdam)(mine, **hp)
    for it in range(4):
        for i, (r, m) in enumerate(zip(ref, mine)):
            if i == 3 and it < 2:  # a parameter without a gradient for the first steps: torch skips it entirely
                r.grad, m.grad = None, None
                continue
            gr = torch.randn(r.shape, generator=g) * (10.0 if it == 1 else 1e-3)  # one step far above the clip, others below
            r.grad, m.grad = gr.clone(), gr.clone().cuda()
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
            o_mine.clip_grad_norm_(max_norm)
        o_ref.step()
        o_mine.step()
        if max_norm > 0:
            n = o_mine.last_grad_norm().cpu()
            assert abs(n[0].item() - total.item()) <= 1e-5 * total.item()
        for r, m in zip(ref, mine):
            assert torch.allclose(m.detach().cpu(), r.detach(), rtol=2e-6, atol=2e-7), (kind, it, r.shape)
    # torch's state layout: checkpoints interchange
    sd = o_mine.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    o_ref.load_state_dict(sd)
    assert torch.allclose(o_ref.state[ref[0]]["exp_avg"].cpu(), o_mine.state[mine[0]]["exp_avg"].cpu())


def test_clip_adam_rejects_cpu_parameters():
    from refid_b200 import optim
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        optim.ClipAdamW([p]).step()
