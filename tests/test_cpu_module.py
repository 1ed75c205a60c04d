"""CPU: host logic of the drop-in module, the C-ABI surface, and the bf16-emulating oracle."""
import ctypes
import os
import re

import pytest
import torch

import flat_util
import paramgen
from oracle import refid_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from refid_b200 import build
    return ctypes.CDLL(build.build())


def test_c_abi_exports_every_declared_symbol():
    L = _lib()
    hdr = open(os.path.join(ROOT, "include", "refid_b200.h")).read()
    names = set(re.findall(r"\b(refid_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/refid_b200.h but not exported"


def test_state_dict_matches_reference_inventory():
    from refid_b200.arch import FinalBidirectionAttenfusion
    for ic, ec in ((26, 2), (6, 2), (3, 5)):
        net = FinalBidirectionAttenfusion(img_chn=ic, ev_chn=ec, num_encoders=3, base_num_channels=32, num_block=1,
                                          num_residual_blocks=2)
        sd, shapes = net.state_dict(), O.param_shapes(ic, ec)
        assert set(sd) == set(shapes)
        assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
        net.load_state_dict(paramgen.make_params(shapes), strict=True)


def test_unsupported_configurations_are_refused():
    from refid_b200.arch import FinalBidirectionAttenfusion
    with pytest.raises(NotImplementedError):
        FinalBidirectionAttenfusion(img_chn=6, ev_chn=2)  # reference defaults: num_encoders=4, num_block=3
    with pytest.raises(NotImplementedError):
        FinalBidirectionAttenfusion(img_chn=6, ev_chn=2, num_encoders=3, num_block=1, skip_type="concat")


def test_flat_vector_layout_and_folds():
    """The flat 'gradient layout' vector: conv weights as [tap][Cin][Cout], folds applied, entries where the engine says."""
    from refid_b200.arch import FinalBidirectionAttenfusion
    net = FinalBidirectionAttenfusion(img_chn=6, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1)
    P = paramgen.make_params(O.param_shapes(6, 2))
    net.load_state_dict(P, strict=True)
    eng = net._table_engine()
    ins, table = net._flat_inputs(eng)  # host logic; the gather itself is CUDA (refid_flat_gather, tested on the GPU)
    flat = flat_util.assemble(eng.flat_floats, ins, table)
    assert flat.numel() == eng.flat_floats and flat.requires_grad
    ent = {e["key"]: e for e in eng.entries}

    def w_of(key):
        e = ent[key]
        return flat[e["w_off"]:e["w_off"] + e["taps"] * e["R"] * e["Cc"]].view(e["taps"], e["R"], e["Cc"]), e

    g, e = w_of("resblocks.0.conv1")
    assert torch.equal(g[4, 7, 9], P["resblocks.0.conv1.weight"][9, 7, 1, 1])
    g, e = w_of("decoders.1.transposed_conv2d")
    assert torch.equal(g[2, 5, 11], P["decoders.1.transposed_conv2d.weight"][11, 5, 1, 0])
    g, e = w_of("head")  # x-unrolled 5x5: k = kx*Cin + c
    assert e["R"] == 32 and torch.equal(g[3, 4 * 2 + 1, 6], P["head.conv2d.weight"][6, 1, 3, 4]) and g[:, 10:].abs().max() == 0
    a = "encoders_forward.1.atten_fuse"
    g, e = w_of(a + ".conv3")
    assert torch.allclose(g[0, 5, 9], P[a + ".conv3.weight"][9, 5, 0, 0] * P[a + ".beta"][0, 9, 0, 0])
    g, e = w_of(a + ".conv5s")
    assert torch.allclose(g[0, 64 + 3, 17], P[a + ".conv5.weight"][17, 3, 0, 0] * P[a + ".gamma"][0, 17, 0, 0])
    assert torch.equal(g[0, 3, 17], P[a + ".conv_y_side.weight"][17, 3, 0, 0])
    bias = flat[e["b_off"]:e["b_off"] + e["nbias"]]
    assert torch.allclose(bias, P[a + ".conv_y_side.bias"] + P[a + ".conv5.bias"] * P[a + ".gamma"].view(-1))
    # the flat vector is differentiable back to every live parameter
    flat.sum().backward()
    dead = set(O.dead_params(O.param_shapes(6, 2)))
    for n, p in net.named_parameters():
        assert (p.grad is None) == (n in dead), n


def test_workspace_accounting():
    from refid_b200 import engine
    eng = engine.Engine(26, 2)
    small = eng.workspace_bytes(1, 2, 32, 32, True)
    big = eng.workspace_bytes(8, 23, 256, 256, True)
    assert 0 < small < 64 << 20
    assert 20 << 30 < big < 120 << 30  # must fit the 180 GB of one B200 with room for inputs
    assert eng.workspace_bytes(8, 23, 256, 256, False) < big
    with pytest.raises(RuntimeError):
        eng.workspace_bytes(1, 2, 30, 32, True)  # H not a multiple of 8: reported, not silently handled


def test_time_chunked_plans_build_for_every_chunk_size():
    """Dry plans (no GPU) of the level-major, time-chunked training schedule: every chunk size plans -- including partial
    last chunks, one step per chunk and chunks longer than T --, the step-major schedule stays available, and the extra
    all-T gradient buffers keep the workspace within 10 % of the step-major plan."""
    from refid_b200 import engine
    eng = engine.Engine(26, 2)
    eng.set_option("tchunk", 0)
    base = eng.workspace_bytes(8, 23, 256, 256, True)
    small0 = eng.workspace_bytes(2, 5, 64, 64, True)
    for k in (1, 2, 3, 5, 8, 23, 64):
        eng.set_option("tchunk", k)
        assert 0.9 * base < eng.workspace_bytes(8, 23, 256, 256, True) < 1.1 * base, k
        assert 0.8 * small0 < eng.workspace_bytes(2, 5, 64, 64, True) < 1.25 * small0, k
        assert eng.workspace_bytes(1, 1, 32, 32, True) > 0
    with pytest.raises(RuntimeError):
        eng.set_option("tchunk", 65)
    # forward-only plans do not depend on the option (they keep the step-major order and O(1)-in-T workspace)
    eng.set_option("tchunk", 0)
    a = eng.workspace_bytes(1, 15, 720, 1280, False)
    eng.set_option("tchunk", 8)
    assert eng.workspace_bytes(1, 15, 720, 1280, False) == a


def test_bf16_emulating_oracle_is_a_restatement_of_the_same_network():
    from oracle import refid_oracle_bf16 as OB
    P = paramgen.make_params(O.param_shapes(6, 2))
    x, ev, _ = paramgen.make_inputs(1, 2, 32, 32, 6, 2)
    with torch.no_grad():
        a, b = O.forward(P, x, ev), OB.forward(P, x, ev)
    assert (a - b).abs().max().item() < 2e-2


def _sync_worker(rank, world, port, q):
    import torch.distributed as dist
    from refid_b200.arch import sync_flat_grad
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.arange(8, dtype=torch.float32) * (rank + 1)
    sync_flat_grad(g, dist.group.WORLD)
    q.put((rank, g.tolist()))
    dist.destroy_process_group()


def test_flat_gradient_all_reduce_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    want = (torch.arange(8, dtype=torch.float32) * 1.5).tolist()
    assert res[0] == want and res[1] == want


def _bcast_worker(rank, world, port, q):
    import torch.distributed as dist
    from refid_b200.arch import FinalBidirectionAttenfusion
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(10 + rank)  # the reference seeds rank r with manual_seed + r (train.py:56): different initial weights
    net = FinalBidirectionAttenfusion(img_chn=6, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1)
    before = float(net.pred.conv2d.weight.double().sum())
    net.grad_sync_group = dist.group.WORLD  # must broadcast rank 0's parameters, as DDP's constructor does
    after = [float(p.double().sum()) for p in net.parameters()]
    q.put((rank, before, after))
    dist.destroy_process_group()


def test_grad_sync_group_broadcasts_rank0_parameters_gloo():
    """the no-DDP data-parallel path must start every replica from rank 0's weights."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = {r: (b, a) for r, b, a in (q.get(timeout=180) for _ in range(2))}
    [p.join(timeout=60) for p in procs]
    assert res[0][0] != res[1][0]      # the ranks really started from different weights
    assert res[0][1] == res[1][1]      # ... and hold rank 0's afterwards
    torch.manual_seed(10)
    from refid_b200.arch import FinalBidirectionAttenfusion
    ref = FinalBidirectionAttenfusion(img_chn=6, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1)
    assert res[1][1] == [float(p.double().sum()) for p in ref.parameters()]


def test_forward_only_plan_is_constant_in_T_and_options_exist():
    from refid_b200 import engine
    eng = engine.Engine(6, 2)
    a, b = eng.workspace_bytes(1, 4, 256, 256, False), eng.workspace_bytes(1, 16, 256, 256, False)
    # only the level-0 in-conv outputs (128 channels, all T) grow with T on a forward-only plan
    per_step = 256 * 256 * 128 * 2
    assert b - a <= 12 * per_step + (1 << 20), (a, b)
    assert eng.workspace_bytes(1, 15, 720, 1280, False) < 8 << 30  # BASELINE.json configs[4] (r1: 50.8 GiB)
    eng.set_option("graphs", 0)
    eng.set_option("infer_fp16", 0)
    with pytest.raises(RuntimeError, match="unknown option"):
        eng.set_option("bogus", 1)
    assert eng.graph_stats() == {"captures": 0, "replays": 0, "eager": 0, "failures": 0}
