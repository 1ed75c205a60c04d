"""GPU: whole-network parity of the CUDA engine (through the C ABI, behind the reference's module interface) against
the golden vectors of the unmodified reference and against the CPU oracles.

Tolerances (BASELINE.json north_star, bf16 path): outputs within 2e-2 max-abs of the reference.  Gradients of a
bf16-storage evaluation of this 100+-layer recurrent net differ from fp32 by several % per tensor (measured with the
bf16-emulating CPU oracle), so gradients are checked (a) by norm and direction against the fp32 oracle and (b) to be no
further from fp32 than an exact bf16-storage evaluation is (oracle/refid_oracle_bf16.py)."""
import pytest
import torch

import golden_util
import paramgen

pytestmark = pytest.mark.gpu


def _net(ic, ec, P, **kw):
    from refid_b200.arch import FinalBidirectionAttenfusion
    net = FinalBidirectionAttenfusion(img_chn=ic, ev_chn=ec, num_encoders=3, base_num_channels=32, num_block=1,
                                      num_residual_blocks=2, **kw)
    net.load_state_dict(P, strict=True)
    return net.cuda()


def _no_abort():
    from refid_b200 import _lib
    torch.cuda.synchronize()
    assert _lib.abort_flag() == 0, "a kernel hit its bounded mbarrier wait"


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("case", list(golden_util.CASES))
def test_forward_backward_vs_reference_golden(case):
    from oracle import refid_oracle as O
    B, T, H, W, ic, ec, x5d = golden_util.CASES[case]
    gold = golden_util.load(case)
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=x5d)
    net = _net(ic, ec, P)
    out = net(x=x.cuda(), event=ev.cuda())
    assert out.shape == gold["out"].shape and out.dtype == torch.float32 and out.is_contiguous()
    err = (out.detach().cpu() - gold["out"]).abs().max().item()
    assert err < 2e-2, f"output max-abs error {err} vs the reference (bf16 tolerance 2e-2)"
    loss = torch.sqrt((out - gt.cuda()) ** 2 + 1e-12).mean()
    assert abs(loss.item() - gold["loss"]) < 2e-3
    loss.backward()
    _no_abort()
    grads = dict(net.named_parameters())
    for n in gold["dead"]:  # parameters the reference never uses get no gradient at all
        assert grads[n].grad is None or grads[n].grad.abs().max().item() == 0.0, n
    bad = []
    for n in gold["names"]:
        if n in gold["dead"]:
            continue
        g = grads[n].grad
        assert g is not None and torch.isfinite(g).all(), n
        mine, ref = g.double().norm().item(), gold["grad_norm"][n]
        if abs(mine - ref) > 0.08 * ref:
            bad.append((n, mine, ref))
    assert not bad, bad[:5]
    # the 32 sampled gradient ELEMENTS per parameter the golden file stores (taken from the unmodified reference): every
    # one within 0.5 x the tensor's RMS gradient, and the median parameter within 0.1 x (bf16 storage of a 100+-layer
    # recurrent net; the fp32 oracle reproduces the same samples to 1e-5, tests/test_oracle_golden.py)
    g_all = {n: p.grad if p.grad is not None else torch.zeros_like(p) for n, p in grads.items()}
    assert not golden_util.check_grads(g_all, gold, rtol_norm=0.08, atol_rel_samples=0.5)
    ratios = golden_util.sample_error_ratios(g_all, gold)
    med = sorted(ratios.values())[len(ratios) // 2]
    print(f"[{case}] sampled-gradient error / tensor RMS: median {med:.3f}, worst {max(ratios.values()):.3f}")
    assert med < 0.1, med


def test_gradients_as_accurate_as_exact_bf16_storage():
    """Fixed cotangent (no sign(pred-gt) discontinuity).  Per parameter: cosine to the fp32 oracle >= 0.95 and
    rel-L2 error <= 1.6 x the error of the bf16-emulating CPU oracle + 0.03."""
    from oracle import refid_oracle as O
    from oracle import refid_oracle_bf16 as OB
    case = "blurry_t3_32"
    B, T, H, W, ic, ec, x5d = golden_util.CASES[case]
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, _ = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=x5d)
    cot = (torch.randn(B, T, 3, H, W, generator=torch.Generator().manual_seed(7)) / (B * T * 3 * H * W)).bfloat16().float()
    Q = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
    (O.forward(Q, x, ev) * cot).sum().backward()
    _, gb = OB.vjp(P, x, ev, cot)
    net = _net(ic, ec, P)
    (net(x=x.cuda(), event=ev.cuda()) * cot.cuda()).sum().backward()
    _no_abort()
    bad = []
    for n, p in net.named_parameters():
        ref = Q[n].grad
        if ref is None:
            continue
        mine = p.grad.detach().cpu()
        cos = ((mine.double() * ref.double()).sum() / (mine.double().norm() * ref.double().norm()).clamp_min(1e-30)).item()
        e_mine, e_bf16 = _rel(mine, ref), _rel(gb[n], ref)
        if cos < 0.95 or e_mine > 1.6 * e_bf16 + 0.03:
            bad.append((n, cos, e_mine, e_bf16))
    assert not bad, bad[:8]


def test_multi_tile_shape_and_eval_mode():
    """128x96, B=2: several spatial tiles per level, ragged tile edges.  A no_grad forward with infer_dtype='bf16' equals
    the training forward bit for bit (same kernels, recycled buffers); the default no_grad forward stores fp16 and is an
    order of magnitude closer to the oracle."""
    from oracle import refid_oracle as O
    B, T, H, W, ic, ec = 2, 2, 96, 128, 6, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=True)
    with torch.no_grad():
        ref = O.forward(P, x, ev)
    net = _net(ic, ec, P, infer_dtype="bf16")
    out = net(x=x.cuda(), event=ev.cuda())
    net.eval()
    with torch.no_grad():
        out_eval = net(x=x.cuda(), event=ev.cuda())
    net16 = _net(ic, ec, P).eval()
    with torch.no_grad():
        out16 = net16(x=x.cuda(), event=ev.cuda())
    _no_abort()
    e_bf16 = (out.detach().cpu() - ref).abs().max().item()
    e_fp16 = (out16.cpu() - ref).abs().max().item()
    print(f"max-abs error vs the fp32 oracle: bf16 plan {e_bf16:.3e}, fp16 forward-only plan {e_fp16:.3e}")
    assert e_bf16 < 2e-2
    assert torch.equal(out.detach(), out_eval)
    assert e_fp16 < 2e-3 and e_fp16 < 0.5 * e_bf16


def _psnr_case(T, H, W, seed):
    """Inputs, parameters with `pred`'s bias moved to mid-range (outputs inside [0,1], so the uint8 quantisation of
    tensor2img is live), oracle output, and a ground truth = oracle output + N(0, sigma) noise at ~36 dB."""
    from oracle import refid_oracle as O
    ic, ec = 26, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=seed)
    P["pred.conv2d.bias"] = P["pred.conv2d.bias"] + 0.5
    x, ev, _ = paramgen.make_inputs(1, T, H, W, ic, ec)
    with torch.no_grad():
        ref = O.forward(P, x, ev)
    g = torch.Generator().manual_seed(99)
    gt = (ref.clamp(0, 1) + 10 ** (-36 / 20) * torch.randn(ref.shape, generator=g)).clamp(0, 1)
    return P, x, ev, ref, gt


def test_headline_sequence_length_T23_parity_and_psnr():
    """T = 23 (the benchmark's sequence length; VERDICT r1: no GPU test went beyond T = 4), 64x64.
    (a) training plan (bf16): output within the north-star's 2e-2 of the fp32 oracle over all 23 frames;
    (b) forward-only plan (fp16 storage): within 2e-3, and PSNR -- through the restated tensor2img / calculate_psnr
        (basicsr/utils/img_util.py:90-117, basicsr/metrics/psnr_ssim.py:47-61) AND through the GPU metric kernel --
        within 0.01 dB of the oracle's PSNR on EVERY frame against a ~36 dB ground truth (a sensitive check: an
        uncorrelated output error of 7.6e-4 RMS moves a 36 dB PSNR by 0.01 dB)."""
    from oracle import refid_oracle as O
    from refid_b200 import metrics
    T, H, W = 23, 64, 64
    P, x, ev, ref, gt = _psnr_case(T, H, W, seed=0)
    net = _net(26, 2, P)
    out_train = net(x=x.cuda(), event=ev.cuda()).detach().cpu()
    net.eval()
    with torch.no_grad():
        out = net(x=x.cuda(), event=ev.cuda())
    _no_abort()
    e_train = (out_train - ref).abs().max().item()
    e_inf = (out.cpu() - ref).abs().max().item()
    rms_inf = (out.cpu() - ref).pow(2).mean().sqrt().item()
    print(f"T=23 64x64 max-abs error vs fp32 oracle: bf16 training plan {e_train:.3e}; fp16 forward-only {e_inf:.3e} (rms {rms_inf:.3e})")
    assert e_train < 2e-2 and e_inf < 2e-3
    # context for the north-star's fp32 bar (1e-3): the same fp32 oracle evaluated through PyTorch-CUDA on this GPU against its
    # CPU evaluation -- with PyTorch's default (cuDNN may run fp32 convolutions in TF32: torch.backends.cudnn.allow_tf32 = True,
    # which the reference never changes) and with TF32 off.  Measured: 9.8e-4 and 6e-6 -- the reference's own CUDA path is as
    # far from its fp32 CPU result as the fp16 forward-only plan is; the network is not chaotic at T = 23.
    Pc = {k: v.cuda() for k, v in P.items()}
    saved = torch.backends.cudnn.allow_tf32
    errs = {}
    try:
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.no_grad():
                rc = O.forward(Pc, x.cuda(), ev.cuda()).cpu()
            errs[tf32] = ((rc - ref).abs().max().item(), (rc - ref).pow(2).mean().sqrt().item())
    finally:
        torch.backends.cudnn.allow_tf32 = saved
    print(f"T=23 64x64: the reference's ops through PyTorch-CUDA vs their CPU fp32 evaluation: cuDNN TF32 allowed (PyTorch "
          f"default) max-abs {errs[True][0]:.3e} rms {errs[True][1]:.3e}; TF32 off max-abs {errs[False][0]:.3e} rms {errs[False][1]:.3e}")
    assert errs[False][0] < 1e-4
    p_gpu = metrics.psnr_frames(out[0], gt[0].cuda(), 0)
    p_ref_gpu = metrics.psnr_frames(ref[0].cuda(), gt[0].cuda(), 0)
    worst = 0.0
    for t in range(T):
        g8 = O.tensor2img_uint8(gt[0, t])
        p_ref = O.psnr_uint8(O.tensor2img_uint8(ref[0, t]), g8)
        p_mine = O.psnr_uint8(O.tensor2img_uint8(out[0, t].cpu()), g8)
        assert 30.0 < p_ref < 42.0, p_ref  # the check is only sensitive around the reference's published PSNR range
        worst = max(worst, abs(p_ref - p_mine))
        assert abs(float(p_gpu[t]) - p_mine) < 1e-9 and abs(float(p_ref_gpu[t]) - p_ref) < 1e-9
    print(f"T=23 worst per-frame |dPSNR| vs the oracle at ~36 dB: {worst:.5f} dB")
    assert worst < 0.01, worst


def test_benchmark_crop_size_forward_parity_256():
    """The benchmark's spatial size and sequence length at B = 1 (cfg2's per-sample shape), forward only: fp16 plan
    within 2e-3 of the fp32 oracle, PSNR at ~36 dB within 0.01 dB on every frame."""
    from oracle import refid_oracle as O
    T, H, W = 23, 256, 256
    P, x, ev, ref, gt = _psnr_case(T, H, W, seed=1)
    net = _net(26, 2, P).eval()
    with torch.no_grad():
        out = net(x=x.cuda(), event=ev.cuda()).cpu()
    _no_abort()
    err = (out - ref).abs().max().item()
    print(f"T=23 256x256 fp16 forward-only max-abs error {err:.3e}, rms {(out - ref).pow(2).mean().sqrt().item():.3e}")
    assert err < 2e-3
    worst = 0.0
    for t in range(T):
        g8 = O.tensor2img_uint8(gt[0, t])
        worst = max(worst, abs(O.psnr_uint8(O.tensor2img_uint8(ref[0, t]), g8) - O.psnr_uint8(O.tensor2img_uint8(out[0, t]), g8)))
    print(f"256x256 worst per-frame |dPSNR| {worst:.5f} dB")
    assert worst < 0.01, worst


def test_cuda_graph_replay_is_bit_identical_and_used():
    """Forward and backward launch lists replayed as CUDA graphs from the second sighting of the same tensors on: same
    bits as plain launches, and the graphs are really used (3 465 launches per step from the host)."""
    from oracle import refid_oracle as O
    B, T, H, W, ic, ec = 1, 3, 64, 64, 6, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = [t.cuda() for t in paramgen.make_inputs(B, T, H, W, ic, ec, x5d=True)]
    cot = torch.randn(B, T, 3, H, W, device="cuda") / (B * T * 3 * H * W)
    res = {}
    for graphs in (0, 1):
        net = _net(ic, ec, P)
        outs = []
        out_buf = None
        for it in range(4):
            for p in net.parameters():
                p.grad = None
            out = net(x=x, event=ev)
            if it == 0:
                st = next(iter(net._states.values()))
                st["engine"].set_option("graphs", graphs)
            out.backward(cot)
            outs.append(out.detach().clone())
            del out
        torch.cuda.synchronize()
        stats = st["engine"].graph_stats()
        res[graphs] = (outs, {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}, stats)
    _no_abort()
    assert res[0][2]["replays"] == 0 and res[0][2]["captures"] == 0
    assert res[1][2]["failures"] == 0 and res[1][2]["captures"] >= 1 and res[1][2]["replays"] >= 2, res[1][2]
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    for n in res[0][1]:
        a, b = res[0][1][n], res[1][1][n]
        assert (a - b).abs().max().item() <= 1e-3 * max(1.0, b.abs().max().item()), n  # fp32 atomics reorder


def test_forward_only_weight_cache_follows_parameter_updates():
    """The packed weights of a forward-only plan are rebuilt only when a parameter changed (version counters), and the
    fused optimizer's raw-pointer update counts as a change."""
    from oracle import refid_oracle as O
    from refid_b200 import optim
    B, T, H, W, ic, ec = 1, 2, 32, 32, 6, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = [t.cuda() for t in paramgen.make_inputs(B, T, H, W, ic, ec, x5d=True)]
    net = _net(ic, ec, P)
    with torch.no_grad():
        a = net(x=x, event=ev).clone()
        key0 = next(iter(net._states.values()))["packed_key"]
        b = net(x=x, event=ev).clone()
        assert next(iter(net._states.values()))["packed_key"] == key0 and torch.equal(a, b)
    opt = optim.ClipAdamW([p for p in net.parameters()], lr=1e-2)
    out = net(x=x, event=ev)
    (out - gt).abs().mean().backward()
    opt.step()
    with torch.no_grad():
        c = net(x=x, event=ev)
    _no_abort()
    assert not torch.equal(a, c), "forward-only plan kept stale packed weights after an optimizer step"


def test_dependent_launch_does_not_change_results():
    """Programmatic dependent launch only moves kernel set-up ahead of the predecessor's completion: the forward output is
    bit-identical with it on and off (ragged 72x104 frames, so depthwise-conv and halo tiles straddle the image edge),
    and the parameter gradients agree to float-atomic reordering."""
    from oracle import refid_oracle as O
    from refid_b200 import _lib
    B, T, H, W, ic, ec = 2, 2, 72, 104, 6, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=3)
    x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=True)
    outs, grads = [], []
    prev = _lib.set_pdl(True)
    try:
        for enable in (True, False):
            _lib.set_pdl(enable)
            net = _net(ic, ec, P)
            out = net(x=x.cuda(), event=ev.cuda())
            (out.float() - gt.cuda()).abs().mean().backward()
            torch.cuda.synchronize()
            outs.append(out.detach().clone())
            grads.append({k: v.grad.detach().clone() for k, v in net.named_parameters() if v.grad is not None})
    finally:
        _lib.set_pdl(prev)
    _no_abort()
    assert torch.equal(outs[0], outs[1])
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 0
    for k in grads[0]:
        a, b = grads[0][k], grads[1][k]
        assert (a - b).abs().max().item() <= 1e-3 * max(1.0, b.abs().max().item()), k


def test_no_cpu_path():
    from refid_b200.arch import FinalBidirectionAttenfusion
    net = FinalBidirectionAttenfusion(img_chn=6, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1)
    with pytest.raises(RuntimeError):
        net(x=torch.rand(1, 6, 32, 32), event=torch.rand(1, 2, 2, 32, 32))


def test_optimize_parameters_and_test_callers_match_oracle():
    """The two callers of the path (reference twoImage_event_recurrent_model.py:273-330) through the option-file ->
    define_network boundary: the loss of one optimize_parameters step equals the oracle's, the AdamW update (after
    the 0.01 global-norm clip) points the same way element-wise, never-used parameters only see weight decay, and
    test() with max_minibatch chunking equals an un-chunked forward."""
    from oracle import refid_oracle as O
    from refid_b200 import recurrent_model
    B, T, H, W, ic, ec = 2, 2, 32, 32, 6, 2
    opt = {"network_g": {"type": "FinalBidirectionAttenfusion", "img_chn": ic, "ev_chn": ec, "num_encoders": 3,
                         "base_num_channels": 32, "num_block": 1, "num_residual_blocks": 2},
           "train": {"optim_g": {"type": "AdamW", "lr": 2e-4, "weight_decay": 1e-4, "betas": [0.9, 0.99]}},
           "val": {"max_minibatch": 1}}
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec)
    m = recurrent_model.TwoImageEventRecurrentRestorationModel(opt, device="cuda")
    m.net_g.load_state_dict(P, strict=True)
    m.feed_data({"lq": x, "voxel": ev, "gt": gt})
    full = m.net_g(x=m.lq, event=m.voxel).detach()
    # chunks of 1 sample == whole batch (no cross-sample coupling); not bit-equal: EGACA's global pool is
    # accumulated with fp32 atomics whose order depends on the launch geometry
    assert (m.test() - full).abs().max().item() < 1e-2  # a flipped bf16 rounding propagates to ~1e-3 at the output
    l_pix = m.optimize_parameters(1)
    _no_abort()
    # oracle: same step on CPU
    Q = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    loss = O.charbonnier(O.forward(Q, x, ev), gt) + 0 * sum(p.sum() for p in Q.values())
    loss.backward()
    torch.nn.utils.clip_grad_norm_(list(Q.values()), 0.01)
    optim = torch.optim.AdamW(list(Q.values()), lr=2e-4, weight_decay=1e-4, betas=(0.9, 0.99))
    optim.step()
    assert abs(l_pix.item() - loss.item()) < 2e-3
    dead = set(O.dead_params(O.param_shapes(ic, ec)))
    agree, total = 0.0, 0
    for n, p in m.net_g.named_parameters():
        d_mine = (p.detach().cpu() - P[n]).double()
        d_ref = (Q[n].detach() - P[n]).double()
        if n in dead:  # zero gradient => pure weight decay, identical on both sides
            assert torch.allclose(d_mine, d_ref, atol=1e-9), n
            continue
        if p.numel() >= 1024:
            agree += (torch.sign(d_mine) == torch.sign(d_ref)).double().sum().item()
            total += p.numel()
    assert agree / total > 0.9, agree / total


def test_full_resolution_inference_shape():
    """BASELINE.json configs[4] geometry (1280x720 frame, forward only, eval/no_grad), two event slices to keep the CPU
    oracle at seconds: 160 x 45 pixel tiles per level-0 conv, ragged tiles at the coarser levels (90 / 45 / 22.5 rows)."""
    from oracle import refid_oracle as O
    B, T, H, W, ic, ec = 1, 2, 720, 1280, 6, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, _ = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=True)
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    with torch.no_grad():
        ref = O.forward(P, x, ev)
    net = _net(ic, ec, P).eval()
    with torch.no_grad():
        out = net(x=x.cuda(), event=ev.cuda())
    _no_abort()
    assert out.shape == (B, T, 3, H, W)
    assert (out.cpu() - ref).abs().max().item() < 2e-2


def test_highrev_crop_training_step():
    """BASELINE.json configs[3] geometry (512x512 crops, img_chn 26) at B=1, T=3: loss and gradient norms against the
    fp32 oracle (norms within 8 %, as for the golden cases)."""
    from oracle import refid_oracle as O
    B, T, H, W, ic, ec = 1, 3, 512, 512, 26, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec)
    ref_out, ref_loss, ref_g = O.loss_and_grads(P, x, ev, gt)
    net = _net(ic, ec, P)
    out = net(x=x.cuda(), event=ev.cuda())
    loss = torch.sqrt((out - gt.cuda()) ** 2 + 1e-12).mean()
    loss.backward()
    _no_abort()
    assert (out.detach().cpu() - ref_out).abs().max().item() < 2e-2
    assert abs(loss.item() - ref_loss.item()) < 2e-3
    bad = []
    for n, p in net.named_parameters():
        if ref_g.get(n) is None or p.grad is None:
            continue
        r = ref_g[n].double().norm().item()
        if r > 0 and abs(p.grad.double().norm().item() - r) > 0.08 * r:
            bad.append((n, p.grad.double().norm().item(), r))
    assert not bad, bad[:5]


def test_flat_parameter_gather_and_scatter_kernels_are_exact():
    """refid_flat_gather / refid_flat_scatter (parameter tensors <-> the engine's flat vector, one launch each way) against
    the torch restatement of the rule (tests/flat_util.py): pure data movement, so bit-exact, forward and backward."""
    import flat_util
    from oracle import refid_oracle as O
    ic, ec = 26, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=3)
    net = _net(ic, ec, P)
    eng = net._table_engine()
    flat = net._flat(eng)
    ins, table = net._flat_inputs(eng)
    want = flat_util.assemble(eng.flat_floats, ins, table)
    assert flat.shape == want.shape and torch.equal(flat.detach(), want.detach())
    g = torch.randn(eng.flat_floats, device="cuda")
    mine = torch.autograd.grad(flat, list(net.parameters()), g, allow_unused=True)
    ref = torch.autograd.grad(want, list(net.parameters()), g, allow_unused=True)
    for (n, _), a, b in zip(net.named_parameters(), mine, ref):
        assert (a is None) == (b is None), n
        if a is not None:
            assert torch.equal(a, b), n
    _no_abort()


def test_time_chunked_schedule_matches_step_major():
    """Training plans run level by level with the recurrence-free ops once per chunk of time steps (engine option "tchunk").
    Against the step-major order of the reference's loop (tchunk = 0): the forward output is bit-identical (same kernels on
    the same per-image data, only batched differently), gradients agree to fp32 summation order -- for chunk sizes that
    divide T, that leave a partial chunk, and for one step per chunk; the chunked plan launches far fewer kernels."""
    from oracle import refid_oracle as O
    B, T, H, W, ic, ec = 2, 5, 64, 64, 26, 2
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=5)
    x, ev, gt = [t.cuda() for t in paramgen.make_inputs(B, T, H, W, ic, ec)]
    res = {}
    for tc in (0, 1, 2, 5, 8):
        net = _net(ic, ec, P)
        net.train_tchunk = tc
        out = net(x=x, event=ev)
        torch.sqrt((out - gt) ** 2 + 1e-12).mean().backward()
        _no_abort()
        eng = next(iter(net._states.values()))["engine"]
        res[tc] = (out.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None},
                   sum(eng.num_launches()))
        net.release_buffers()
    out0, g0, n0 = res[0]
    for tc in (1, 2, 5, 8):
        out, g, n = res[tc]
        assert torch.equal(out, out0), f"tchunk={tc}: forward output differs from the step-major schedule"
        assert set(g) == set(g0)
        worst = max(_rel(g[k], g0[k]) for k in g0 if g0[k].abs().max() > 0)
        print(f"tchunk={tc}: {n} launches per step (step-major {n0}); worst per-parameter gradient rel-L2 difference {worst:.2e}")
        assert worst < 2e-2, (tc, worst)  # bf16 gradient buffers: a few accumulations are associated differently
    assert res[5][2] < 0.7 * n0
