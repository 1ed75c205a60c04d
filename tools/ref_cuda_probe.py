"""Context number (not a bench line): the reference's algorithm as plain torch ops (the oracle restatement, which issues
the same ATen/cuDNN calls as the reference module) on the B200, fp32 (TF32 convs, cudnn.benchmark) and bf16 autocast +
channels_last.  Diagnostic only; usage: python tools/ref_cuda_probe.py [B] [T]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import paramgen
from oracle import refid_oracle as O
import bench

B, T = [int(v) for v in (sys.argv[1:3] + ["2", "23"][len(sys.argv) - 1:])]
H = W = 256
torch.backends.cudnn.benchmark = True
res = {}
for mode in ("fp32_tf32", "bf16_autocast"):
    P = {k: v.cuda().requires_grad_(True) for k, v in paramgen.make_params(O.param_shapes(26, 2), seed=0).items()}
    x, ev, gt = bench.make_inputs(B, T, H, W, 26, 2, seed=1234, device="cuda")
    def step():
        for p in P.values():
            p.grad = None
        with torch.autocast("cuda", torch.bfloat16, enabled=(mode == "bf16_autocast")):
            out = O.forward(P, x, ev)
        loss = torch.sqrt((out.float() - gt) ** 2 + 1e-12).mean()
        loss.backward()
    try:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        res[mode] = {"B": B, "T": T, "ms_per_step": dt * 1e3, "frames_per_s": B * T / dt,
                     "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    except Exception as e:  # noqa: BLE001
        res[mode] = {"error": repr(e)[:200]}
    print(mode, res[mode], flush=True)
    del P
    torch.cuda.empty_cache()
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_probe.json"), "w"), indent=1)
