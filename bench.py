#!/usr/bin/env python
"""Benchmark of the REFID hot path (forward + Charbonnier loss + backward of FinalBidirectionAttenfusion).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload gopro_11p1|gopro_7skip|tiny]

One "step" = one training iteration's network work on one batch of synthetic GoPro-shaped input:
forward, Charbonnier loss, backward to every parameter gradient, and (N > 1) the NCCL all-reduce of the flat gradient.
Metric: output frames / second = B*T*N / step time.  Default workload = BASELINE.json configs[1]
(GoPro blurry VFI 11+1: x (8,26,256,256), event (8,23,2,256,256), bf16 compute, per GPU).

`--impl reference` times the reference's algorithm on the host CPU cores (the fp32 oracle port in oracle/, all
threads) on a bounded sample of the same workload.  Prints ONE JSON line on rank 0.
"""
import argparse
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
    "gopro_11p1": (8, 23, 256, 256, 26, 2),   # BASELINE.json configs[1]
    "gopro_7skip": (8, 7, 256, 256, 6, 2),    # configs[2] per GPU
    "highrev_11p3": (2, 25, 512, 512, 26, 2),  # configs[3] per GPU
    "tiny": (1, 3, 64, 64, 26, 2),
}
METRIC = "frames/sec fwd+bwd 256x256 GoPro 11+1"


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


def cpu_reference_run(B, T, H, W, ic, ec, steps, warmup, sample_T=None, sample_hw=None):
    """The fp32 oracle port (oracle/refid_oracle.py) fwd + Charbonnier + bwd on the host cores."""
    import torch
    from oracle import refid_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import paramgen  # deterministic O(1)-scale parameters shared with the parity tests
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Ts = sample_T or T
    Hs, Ws = sample_hw or (H, W)
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = make_inputs(1, Ts, Hs, Ws, ic, ec, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.loss_and_grads(P, x, ev, gt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    # per-frame CPU cost is linear in pixels and in T (same convs per pixel per step)
    fps = Ts / sec * (Hs * Ws) / (H * W)
    sample = f"B=1, T={Ts}, {Hs}x{Ws} crop of the workload, fwd+Charbonnier+bwd, fp32, {cores} threads, {steps} timed steps; " \
             f"frames/s scaled by pixel count to {H}x{W}"
    return fps, sec, cores, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gopro_11p1", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    B, T, H, W, ic, ec = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: x ({B},{ic},{H},{W}) + event ({B},{T},{ec},{H},{W}) per GPU, T={T} output frames",
              "batch_per_gpu": B, "T": T, "H": H, "W": W, "img_chn": ic, "ev_chn": ec,
              "step": "forward + Charbonnier loss + backward to all parameter gradients (+ flat-gradient NCCL all-reduce if N>1)",
              "l2": "per-step working set (tens of GB of saved activations) far exceeds the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        # bounded sample: a 128x128 crop with T=4 event slices keeps one step at a few seconds of CPU time
        fps, sec, cores, sample = cpu_reference_run(B, T, H, W, ic, ec, max(1, args.steps), max(1, min(args.warmup, 1)),
                                                    sample_T=min(T, 4), sample_hw=(min(H, 128), min(W, 128)))
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from refid_b200.arch import FinalBidirectionAttenfusion

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    net = FinalBidirectionAttenfusion(img_chn=ic, ev_chn=ec, num_encoders=3, base_num_channels=32, num_block=1,
                                      num_residual_blocks=2)
    init_params(net)
    net = net.to(dev)
    if world > 1:
        net.grad_sync_group = dist.group.WORLD  # flat-gradient all-reduce (mean) inside the backward
    hx, hev, hgt = make_inputs(B, T, H, W, ic, ec, seed=1234 + rank, pin=True)
    x, ev, gt = hx.to(dev), hev.to(dev), hgt.to(dev)
    h2d = sum(t.numel() * t.element_size() for t in (hx, hev, hgt))

    from refid_b200.losses import CharbonnierLoss
    cri_pix = CharbonnierLoss(loss_weight=1.0, reduction="mean")  # train.pixel_opt of the GoPro option files

    def step(xd, evd, gtd):
        for p in net.parameters():
            p.grad = None
        out = net(x=xd, event=evd)
        loss = cri_pix(out, gtd)  # twoImage_event_recurrent_model.py:284; value + gradient in one CUDA pass (csrc/loss.cu)
        loss.backward()
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

    for _ in range(max(args.warmup, 3)):
        step(x, ev, gt)
    with ClockSampler(local_rank) as clk:
        ms = timed(lambda: step(x, ev, gt), args.steps)
    frames = B * T * world * args.steps
    value = frames / (ms * 1e-3)

    # end to end: every step copies its inputs from pinned host memory and reads the loss back to the host.  As in the
    # reference's CUDAPrefetcher (basicsr/data/prefetch_dataloader.py) the copy of step i+1 runs on a side stream while
    # step i computes; one copy is issued per step inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)

    def h2d_async():
        with torch.cuda.stream(copy_stream):
            t = (hx.to(dev, non_blocking=True), hev.to(dev, non_blocking=True), hgt.to(dev, non_blocking=True))
            e = torch.cuda.Event()
            e.record(copy_stream)
        return t, e

    pending = [h2d_async()]

    def e2e_step():
        (xd, evd, gtd), e = pending.pop()
        cur = torch.cuda.current_stream()
        cur.wait_event(e)
        pending.append(h2d_async())
        loss = step(xd, evd, gtd)
        for t in (xd, evd, gtd):
            t.record_stream(cur)
        return loss.item()

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = frames / (ms_e2e * 1e-3)

    st = next(s for k, s in net._states.items() if k[4])
    nf, nb = st["engine"].num_launches()
    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": (nf + nb + 3) * args.steps,  # plan launches + loss, loss finish, upstream-gradient scale
            "clocks": clk.summary(),
            "samples_per_s": value / T}
    if rank == 0:
        pk = peaks()
        gf = 3.0 * algorithmic_gflop_fwd(T, H, W, ic, ec) * B  # fwd + dgrad + wgrad, per GPU per step
        step_tflops = gf / (ms / args.steps)  # GFLOP / ms = TFLOP/s
        line["step_tflops_algorithmic"] = step_tflops
        line["step_frac_of_bf16_sustained"] = step_tflops / pk["bf16_sustained"]
        if not args.no_profile:
            prof = st["engine"].profile(True)
            torch.cuda.synchronize()
            # dominant kernel: haloconv_kernel on the stride-1 3x3 convs with >= 64 channels (forward + data-gradient)
            conv = {k: prof[k] for k in ("conv3x3_fwd", "conv3x3_dgrad")}
            cms = sum(v["ms"] for v in conv.values())
            cfl = sum(v["flops"] for v in conv.values())
            cn = sum(v["launches"] for v in conv.values())
            tot = sum(v["ms"] for v in prof.values())
            ach = cfl / (cms * 1e-3) / 1e12
            traffic = None
            tp = os.path.join(ROOT, "profiles", "r01_ncu_haloconv.json")
            if os.path.exists(tp):  # dram bytes per launch from the committed `ncu --set full` capture of the same kernel
                caps = [c for c in json.load(open(tp)) if "haloconv" in c.get("kernel", "")]
                if caps:
                    traffic = 1e6 * sum(c["dram_read_MB"] + c["dram_write_MB"] for c in caps) / len(caps)
            line["roofline"] = {"bound": "tensor",
                                "kernel": "haloconv_kernel (stride-1 3x3 implicit-GEMM conv, >= 64 channels, forward + data-gradient, tcgen05)",
                                "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                                "peak_source": pk["source"] + " bf16_tflops_sustained (kernel timed inside a long step)",
                                "traffic": traffic, "traffic_note": "mean dram read+write bytes per captured launch (profiles/r01_ncu_haloconv.json)",
                                "launches_per_step": cn, "avg_launch_ms": cms / max(cn, 1),
                                "share_of_step_device_time": cms / tot,
                                "per_class": {k: {"ms": v["ms"], "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] else 0.0,
                                                  "launches": v["launches"]} for k, v in prof.items()}}
        if not args.no_cpu_baseline and world == 1:
            fps, sec, cores, sample = cpu_reference_run(B, T, H, W, ic, ec, 2, 1, sample_T=min(T, 4),
                                                        sample_hw=(min(H, 128), min(W, 128)))
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
