"""CPU: the drop-in boundary -- option file -> define_network -> module, as the reference's wrappers call it."""
import os
from copy import deepcopy

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_option_files_build_the_backend_through_define_network():
    from refid_b200 import plugin
    from refid_b200.arch import FinalBidirectionAttenfusion
    for rel in ("options/train/GoPro_blurry_11p1_b200.yml", "options/test/GoPro_blurry_11p1_b200.yml"):
        opt = plugin.load_options(os.path.join(ROOT, rel))
        ng = deepcopy(opt["network_g"])
        net = plugin.define_network(ng)
        assert isinstance(net, FinalBidirectionAttenfusion) and "type" not in ng
        assert net.img_chn == 26 and net.ev_chn == 2
        assert len(net.state_dict()) == 183 and sum(p.numel() for p in net.parameters()) == 15928355


def test_lookup_semantics_match_the_reference():
    from refid_b200 import plugin
    mods = plugin.scan_arch_modules()
    assert [m.__name__ for m in mods] == ["refid_b200.archs.refid_b200_arch"]  # exactly one definition is scanned
    with pytest.raises(ValueError, match="is not found"):
        plugin.define_network({"type": "UNetRecurrent", "img_chn": 6})
    with pytest.raises(TypeError):  # unknown kwargs reach the constructor unchanged, as in the reference
        plugin.define_network({"type": "FinalBidirectionAttenfusion", "img_chn": 6, "ev_chn": 2, "bogus": 1})


def test_wrapper_needs_cuda_and_keeps_reference_defaults():
    import torch
    from refid_b200 import plugin, recurrent_model
    opt = plugin.load_options(os.path.join(ROOT, "options/train/GoPro_blurry_11p1_b200.yml"))
    m = recurrent_model.TwoImageEventRecurrentRestorationModel(opt, device="cpu")
    from refid_b200 import optim
    assert isinstance(m.optimizer_g, optim.ClipAdamW) and isinstance(m.optimizer_g, torch.optim.Optimizer)
    g = m.optimizer_g.param_groups[0]
    assert g["lr"] == 2e-4 and g["weight_decay"] == 1e-4 and tuple(g["betas"]) == (0.9, 0.99)
    m.feed_data({"lq": torch.rand(1, 26, 32, 32), "voxel": torch.rand(1, 2, 2, 32, 32), "gt": torch.rand(1, 2, 3, 32, 32)})
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.optimize_parameters(1)
