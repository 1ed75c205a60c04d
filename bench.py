#!/usr/bin/env python
"""Benchmark of the REFID hot path (FinalBidirectionAttenfusion forward / backward on the sm_100a engine).

    python bench.py --gpus N --steps K --warmup W [--impl reference|reference-cuda] [--workload NAME] [--mode train|infer]

Training mode (default): one "step" = one training iteration's network work on one batch of synthetic input (SURVEY.md
8d): forward, Charbonnier loss, backward to every parameter gradient, and (N > 1) the NCCL all-reduce of the flat
gradient.  Inference mode: one step = one no_grad forward.  Metric: output frames / second = B*T*N / step time.
Default workload = BASELINE.json configs[1] (GoPro blurry VFI 11+1: x (8,26,256,256), event (8,23,2,256,256), bf16).

Reference arms (rank 0 only):
  --impl reference       the UNMODIFIED reference module (oracle/_ref staged files, else /root/reference; the oracle port if
                         neither is there) on the host cores, all threads, fp32, on ONE sample of the same workload at
                         full T and resolution (no extrapolation: the reference's per-frame CPU cost does not depend on B);
  --impl reference-cuda  the same unmodified module on the GPU through PyTorch/cuDNN: fp32 with TF32 convolutions and
                         cudnn.benchmark (the reference's train.py setting) and bf16 autocast + channels_last (the stronger
                         baseline), at the largest batch <= the workload's that fits.
Prints ONE JSON line on rank 0.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (B per GPU, T, H, W, img_chn, ev_chn)
    "gopro_11p1": (8, 23, 256, 256, 26, 2),     # BASELINE.json configs[1]  (the metric's configuration)
    "gopro_7skip": (8, 7, 256, 256, 6, 2),      # configs[2] per GPU
    "highrev_11p3": (2, 25, 512, 512, 26, 2),   # configs[3] per GPU
    "fullres_720p": (1, 15, 720, 1280, 6, 2),   # configs[4]: inference only
    "gopro_11p1_b1": (1, 23, 256, 256, 26, 2),  # the reference's own training batch (1 per GPU): launch-bound regime
    "tiny": (1, 3, 64, 64, 26, 2),
}
METRIC = "frames/sec fwd+bwd 256x256 GoPro 11+1"
NET_KW = dict(num_encoders=3, base_num_channels=32, num_block=1, num_residual_blocks=2)


def algorithmic_gflop_fwd(T, H, W, img_chn, ev_chn):
    """Forward GFLOP per sample: s*[(28.99 + 0.10486*img) + T*(163.13 + 0.10486*ev)]."""
    s = H * W / 65536.0
    return s * ((28.99 + 0.10486 * img_chn) + T * (163.13 + 0.10486 * ev_chn))


def make_inputs(B, T, H, W, ic, ec, seed, device="cpu", pin=False):
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, ic, H, W, generator=g)
    if ic == 26:
        idx = list(range(3, 13)) + list(range(16, 26))
        x[:, idx] = (torch.randn(B, len(idx), H, W, generator=g) * (torch.rand(B, len(idx), H, W, generator=g) > 0.8))
    ev = torch.randn(B, T, ec, H, W, generator=g) * (torch.rand(B, T, ec, H, W, generator=g) > 0.8)
    gt = torch.rand(B, T, 3, H, W, generator=g)
    if pin:
        x, ev, gt = x.pin_memory(), ev.pin_memory(), gt.pin_memory()
    if device != "cpu":
        x, ev, gt = x.to(device), ev.to(device), gt.to(device)
    return x, ev, gt


def init_params(net, seed=0):
    """Reference default init (same seed => same weights as the reference constructor) + non-zero beta/gamma
    (zero-initialised gates would switch the attention branch off)."""
    import torch
    g = torch.Generator().manual_seed(seed + 99)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith(".beta") or n.endswith(".gamma"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))


def bench_state_dict(ic, ec):
    """The benchmark's parameters as a CPU state_dict (both arms and the parity check use exactly these)."""
    import torch
    from refid_b200.arch import FinalBidirectionAttenfusion
    torch.manual_seed(0)
    net = FinalBidirectionAttenfusion(img_chn=ic, ev_chn=ec, **NET_KW)
    init_params(net)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.rows, self.stop = [], threading.Event()
        self.index = index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(float(r[0])) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(self.rows[0][1])), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------------------------------
# reference arms (the only places bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def reference_module(ic, ec, state_dict):
    """(module, kind): the unmodified reference network with the benchmark's parameters, or the oracle port."""
    from oracle import ref_loader
    if ref_loader.available():
        with contextlib.redirect_stdout(io.StringIO()):
            net = ref_loader.build(ic, ec)
        net.load_state_dict(state_dict, strict=True)
        return net, "reference", ref_loader.source()
    import torch
    from oracle import refid_oracle as O

    class Port(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.P = torch.nn.ParameterDict({k.replace(".", "/"): torch.nn.Parameter(v.clone()) for k, v in state_dict.items()})

        def forward(self, x, event):
            return O.forward({k.replace("/", "."): v for k, v in self.P.items()}, x, event)

    return Port(), "port", "oracle/refid_oracle.py (reference files not staged)"


def charbonnier(out, gt):
    import torch
    return torch.sqrt((out.float() - gt) ** 2 + 1e-12).mean()  # basicsr/models/losses/losses.py:28-30


def cpu_reference_run(T, H, W, ic, ec, steps, warmup, train=True, state_dict=None, inputs=None):
    """The reference on the host cores: ONE sample of the workload at full T and resolution; fwd + Charbonnier + bwd (or a
    no_grad forward).  Returns (frames/s, s/step, cores, kind, sample text, output of the last step)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = state_dict if state_dict is not None else bench_state_dict(ic, ec)
    net, kind, src = reference_module(ic, ec, sd)
    x, ev, gt = inputs if inputs is not None else make_inputs(1, T, H, W, ic, ec, seed=1234)
    times, out = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if train:
            for p in net.parameters():
                p.grad = None
            out = net(x=x, event=ev)
            charbonnier(out, gt).backward()
        else:
            with torch.no_grad():
                out = net(x=x, event=ev)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    what = "fwd+Charbonnier+bwd" if train else "no_grad forward"
    sample = (f"one sample (B=1) of the workload at full size: T={T}, {H}x{W}, {what}, fp32, {cores} threads, {steps} timed "
              f"step(s) after {warmup} warm-up; {src}; frames/s = T / step time (per-frame CPU cost is independent of B)")
    return T / sec, sec, cores, kind, sample, out.detach()


def cuda_reference_run(B, T, H, W, ic, ec, steps, warmup, train=True, state_dict=None):
    """The unmodified reference module on the GPU through PyTorch/cuDNN.  Two settings: fp32 with TF32 convs and
    cudnn.benchmark (the reference's train.py:139), and bf16 autocast + channels_last.  Batch = the largest power-of-two
    fraction of B that fits (the fp32 path keeps ~0.75 GB of activations per step and sample at 256^2)."""
    import torch
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    sd = state_dict if state_dict is not None else bench_state_dict(ic, ec)
    res = {}
    for mode in ("fp32_tf32", "bf16_autocast_channels_last"):
        b = B
        while b >= 1:
            try:
                net, kind, src = reference_module(ic, ec, sd)
                net = net.cuda()
                if mode != "fp32_tf32":
                    net = net.to(memory_format=torch.channels_last)
                x, ev, gt = make_inputs(b, T, H, W, ic, ec, seed=1234, device="cuda")

                def step():
                    with torch.autocast("cuda", torch.bfloat16, enabled=(mode != "fp32_tf32")):
                        if train:
                            for p in net.parameters():
                                p.grad = None
                            out = net(x=x, event=ev)
                        else:
                            with torch.no_grad():
                                out = net(x=x, event=ev)
                    if train:
                        charbonnier(out, gt).backward()
                for _ in range(max(warmup, 2)):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                res[mode] = {"batch": b, "ms_per_step": ms, "frames_per_s": b * T / (ms * 1e-3), "kind": kind,
                             "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
                break
            except torch.cuda.OutOfMemoryError:
                b //= 2
            finally:
                net = x = ev = gt = None
                torch.cuda.empty_cache()
        if mode not in res:
            res[mode] = {"error": "out of memory at batch 1"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="gopro_11p1", choices=list(WORKLOADS))
    ap.add_argument("--mode", default=None, choices=["train", "infer"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    args = ap.parse_args()
    B, T, H, W, ic, ec = WORKLOADS[args.workload]
    mode = args.mode or ("infer" if args.workload == "fullres_720p" else "train")
    train = mode == "train"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric = METRIC if (args.workload == "gopro_11p1" and train) else \
        f"frames/sec {'fwd+bwd' if train else 'forward (no_grad)'} {args.workload}"
    step_text = ("forward + Charbonnier loss + backward to all parameter gradients (+ flat-gradient NCCL all-reduce if N>1)"
                 if train else "no_grad forward (fp16 storage), output kept on the device")
    config = {"workload": f"{args.workload}: x ({B},{ic},{H},{W}) + event ({B},{T},{ec},{H},{W}) per GPU, T={T} output frames",
              "mode": mode, "batch_per_gpu": B, "T": T, "H": H, "W": W, "img_chn": ic, "ev_chn": ec, "step": step_text,
              "l2": "per-step working set (GBs of activations) far exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        fps, sec, cores, kind, sample, _ = cpu_reference_run(T, H, W, ic, ec, max(1, args.steps), max(0, min(args.warmup, 1)),
                                                            train=train)
        line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    if args.impl == "reference-cuda":
        if rank != 0:
            return
        res = cuda_reference_run(B, T, H, W, ic, ec, max(1, args.steps), args.warmup, train=train)
        ok = {k: v for k, v in res.items() if "frames_per_s" in v}
        best = max(ok, key=lambda k: ok[k]["frames_per_s"]) if ok else None
        line = {"impl": "reference-cuda", "metric": metric, "value": ok[best]["frames_per_s"] if best else None,
                "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ok[best]["ms_per_step"] if best else None, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": best, "data": "synthetic", "config": config, "modes": res,
                "note": "unmodified reference module on the same GPU through PyTorch/cuDNN; value = the faster setting"}
        print(json.dumps(line))
        return

    from refid_b200 import _lib
    from refid_b200.arch import FinalBidirectionAttenfusion
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd = bench_state_dict(ic, ec)
    net = FinalBidirectionAttenfusion(img_chn=ic, ev_chn=ec, **NET_KW)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev)
    if not train:
        net.eval()
    if world > 1 and train:
        net.grad_sync_group = dist.group.WORLD  # flat-gradient all-reduce (mean) inside the backward
    hx, hev, hgt = make_inputs(B, T, H, W, ic, ec, seed=1234 + rank, pin=True)
    x, ev, gt = hx.to(dev), hev.to(dev), hgt.to(dev)
    h2d = sum(t.numel() * t.element_size() for t in ((hx, hev, hgt) if train else (hx, hev)))

    from refid_b200.losses import CharbonnierLoss
    cri_pix = CharbonnierLoss(loss_weight=1.0, reduction="mean")  # train.pixel_opt of the GoPro option files
    last = {}

    def step(xd, evd, gtd):
        if not train:
            with torch.no_grad():
                out = net(x=xd, event=evd)
            last["out"] = out
            return out
        for p in net.parameters():
            p.grad = None
        out = net(x=xd, event=evd)
        loss = cri_pix(out, gtd)  # twoImage_event_recurrent_model.py:284; value + gradient in one CUDA pass (csrc/loss.cu)
        loss.backward()
        last["out"], last["loss"] = out, loss
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def settle(fn, limit=8):
        """Extra UNTIMED steps until two consecutive steps were pure CUDA-graph replays.  The engine captures a graph at the
        second sighting of an (x, event, out) pointer set and PyTorch's allocator alternates between two `out` blocks, so
        after exactly three warm-up steps the first timed step would still pay one capture (~100 ms: +12 ms per step at
        K = 8).  Returns the number of extra steps (reported in `warmup`)."""
        engines = [s["engine"] for s in net._states.values()]
        def counts():
            g = [e.graph_stats() for e in engines]
            return sum(v["captures"] + v["eager"] + v["failures"] for v in g)
        extra = clean = 0
        if world > 1:  # every step contains the gradient all-reduce: all ranks must run the SAME number of steps
            for _ in range(4):
                fn()
            torch.cuda.synchronize()
            return 4
        while clean < 2 and extra < limit:
            before = counts()
            fn()
            torch.cuda.synchronize()
            extra += 1
            clean = clean + 1 if counts() == before else 0
        return extra

    for _ in range(max(args.warmup, 3)):
        step(x, ev, gt)
    warm_extra = settle(lambda: step(x, ev, gt)) if not args.no_graphs else 0
    if args.no_graphs:
        for s in net._states.values():
            s["engine"].set_option("graphs", 0)
    with ClockSampler(local_rank) as clk:
        ms = timed(lambda: step(x, ev, gt), args.steps)
    _lib.raise_if_aborted()
    frames = B * T * world * args.steps
    value = frames / (ms * 1e-3)
    out0 = last["out"][0].detach().float().cpu()  # sample 0 of the last timed step, for the parity record below
    if train:
        loss_value = float(last["loss"].item())
        if not (loss_value == loss_value and abs(loss_value) < 1e6):
            raise SystemExit(f"bench: the training loss is not finite ({loss_value})")
    if not bool(torch.isfinite(last["out"]).all()):
        raise SystemExit("bench: the network output is not finite")

    # end to end: every step copies its inputs from pinned host memory and reads the result (the loss; in inference mode
    # the mean of the output) back to the host.  As in the reference's CUDAPrefetcher (basicsr/data/prefetch_dataloader.py)
    # the copy of step i+1 runs on a side stream while step i computes; one copy is issued per step inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    hosts = (hx, hev, hgt) if train else (hx, hev)
    # two alternating sets of device input buffers: the engine keys its CUDA graphs on the input pointers, and the host now
    # runs ahead of the device (no blocking read per step), so fresh allocations per step would mean a new graph key per step
    NBUF = 2
    dsets = [tuple(torch.empty_like(h, device=dev) for h in hosts) for _ in range(NBUF)]
    consumed = [None] * NBUF  # event: the step that read buffer set j has finished
    issued = [0]

    def h2d_async():
        j = issued[0] % NBUF
        issued[0] += 1
        with torch.cuda.stream(copy_stream):
            if consumed[j] is not None:
                copy_stream.wait_event(consumed[j])
            for dst, h in zip(dsets[j], hosts):
                dst.copy_(h, non_blocking=True)
            e = torch.cuda.Event()
            e.record(copy_stream)
        return dsets[j], e, j

    pending = [h2d_async()]

    # The step's result (the loss; in inference mode the mean of the output) is read back EVERY step by an asynchronous copy
    # into pinned host memory that is consumed one step later -- the way a training loop logs (the reference reduces and
    # reads its losses only at print time, basicsr/train.py) -- so the host enqueues step i+1 while step i computes instead
    # of idling the GPU for the ~5 ms of enqueue after every blocking .item().  All reads are complete (and checked) inside
    # the timed region: the closing event is recorded after the last copy.
    results = torch.zeros(args.steps + 2, dtype=torch.float32).pin_memory()
    count = [0]

    def e2e_step():
        t, e, j = pending.pop()
        cur = torch.cuda.current_stream()
        cur.wait_event(e)
        pending.append(h2d_async())
        r = step(t[0], t[1], t[2] if train else None)
        results[count[0] % results.numel()].copy_((r if train else r.mean()).detach().reshape(()), non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        consumed[j] = done
        count[0] += 1

    e2e_step()
    if not args.no_graphs:
        settle(e2e_step)
    count[0] = 0
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = frames / (ms_e2e * 1e-3)
    if not bool(torch.isfinite(results[:args.steps]).all()):
        raise SystemExit("bench: a step result read back in the end-to-end loop is not finite")

    st = next(s for k, s in net._states.items() if bool(k[4]) == train)
    nf, nb = st["engine"].num_launches()
    gstats = st["engine"].graph_stats()
    line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3) + warm_extra, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if train else "fp16", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "note": "inputs: pinned host -> device on a copy stream, one step ahead; result: async copy to pinned host "
                            "memory every step, consumed one step later; all inside the timed region"},
            # plan launches (+ loss, loss finish, upstream-gradient scale in training); replayed from CUDA graphs
            "gpu_launches": (nf + (nb + 3 if train else 0)) * args.steps,
            "cuda_graphs": gstats,
            "clocks": clk.summary(),
            "samples_per_s": value / T,
            "workspace_gib": st["engine"].workspace_bytes(B, T, H, W, train) / 2 ** 30}
    if rank == 0:
        pk = peaks()
        gf = (3.0 if train else 1.0) * algorithmic_gflop_fwd(T, H, W, ic, ec) * B  # fwd (+ dgrad + wgrad), per GPU per step
        step_tflops = gf / (ms / args.steps)  # GFLOP / ms = TFLOP/s
        line["step_tflops_algorithmic"] = step_tflops
        line["step_frac_of_bf16_sustained"] = step_tflops / pk["bf16_sustained"]
        if not args.no_profile:
            prof = st["engine"].profile(train)
            torch.cuda.synchronize()
            # dominant kernel: haloconv_kernel on the stride-1 3x3 convs with >= 64 channels (forward + data-gradient)
            conv = {k: prof[k] for k in ("conv3x3_fwd", "conv3x3_dgrad")}
            cms = sum(v["ms"] for v in conv.values())
            cfl = sum(v["flops"] for v in conv.values())
            cn = sum(v["launches"] for v in conv.values())
            tot = sum(v["ms"] for v in prof.values())
            ach = cfl / (cms * 1e-3) / 1e12
            traffic = None
            for name in ("r02_ncu_haloconv3x3_lean.json", "r02_ncu_haloconv3x3.json", "r01_ncu_haloconv.json"):
                tp = os.path.join(ROOT, "profiles", name)
                if os.path.exists(tp):  # dram bytes per launch from the committed `ncu --set full` capture of the same kernel
                    caps = [c for c in json.load(open(tp)) if "haloconv" in c.get("kernel", "")]
                    if caps:
                        traffic = 1e6 * sum(c["dram_read_MB"] + c["dram_write_MB"] for c in caps) / len(caps)
                        traffic_src = name
                        break
            line["roofline"] = {"bound": "tensor",
                                "kernel": "haloconv_kernel (stride-1 3x3 implicit-GEMM conv, >= 64 channels, forward + data-gradient, tcgen05)",
                                "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                                "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                                "traffic": traffic,
                                "traffic_note": f"mean dram read+write bytes per captured launch (profiles/{traffic_src})" if traffic else None,
                                "launches_per_step": cn, "avg_launch_ms": cms / max(cn, 1),
                                "share_of_step_device_time": cms / tot,
                                "per_class": {k: {"ms": v["ms"], "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] else 0.0,
                                                  "launches": v["launches"]} for k, v in prof.items()}}
        if not args.no_cpu_baseline and world == 1:
            # the reference on the host cores on sample 0 of this very batch, with these very parameters: the timing is the
            # cpu_baseline, its output is the parity record of the timed configuration (all T frames of sample 0)
            fps, sec, cores, kind, sample, ref_out = cpu_reference_run(
                T, H, W, ic, ec, 1, 0, train=train, state_dict=sd, inputs=(hx[0:1].clone(), hev[0:1].clone(), hgt[0:1].clone()))
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample}
            d = (out0 - ref_out[0]).abs()
            line["parity"] = {"max_abs_err_vs_reference": d.max().item(), "rms_err": d.pow(2).mean().sqrt().item(),
                              "frames": T, "sample": 0, "tolerance": 2e-2 if train else 2e-3,
                              "storage": "bf16" if train else "fp16"}
            if not d.max().item() < line["parity"]["tolerance"]:
                raise SystemExit(f"bench: output of the timed configuration differs from the reference: {line['parity']}")
        if not args.no_ref_cuda and world == 1 and train:
            # context for the north-star's ">= 5x the reference PyTorch-CUDA path": free our workspace first
            net.release_buffers()
            del st
            torch.cuda.empty_cache()
            rc = cuda_reference_run(B, T, H, W, ic, ec, 2, 2, train=True, state_dict=sd)
            line["reference_cuda"] = rc
            best = max((v["frames_per_s"] for v in rc.values() if "frames_per_s" in v), default=None)
            line["ref_cuda_frames_s"] = best
            line["speedup_vs_reference_cuda"] = (value / best) if best else None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
