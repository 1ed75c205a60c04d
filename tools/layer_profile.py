"""Per-launch device times of one training step (CUDA events around every launch of the engine's plan), aggregated per
parameter site and pass.  Usage: python tools/layer_profile.py [B] [T] [H] [W] [tag]"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from refid_b200.arch import FinalBidirectionAttenfusion  # noqa: E402

a = sys.argv[1:]
B, T, H, W = [int(v) for v in (a[:4] + ["8", "4", "256", "256"][len(a[:4]):])]
tag = a[4] if len(a) > 4 else "layers"
torch.manual_seed(0)
net = FinalBidirectionAttenfusion(img_chn=26, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1)
bench.init_params(net)
net = net.cuda()
x, ev, gt = bench.make_inputs(B, T, H, W, 26, 2, seed=1234, device="cuda")
for it in range(3):
    for p in net.parameters():
        p.grad = None
    out = net(x=x, event=ev)
    loss = torch.sqrt((out - gt) ** 2 + 1e-12).mean()
    loss.backward()
torch.cuda.synchronize()
st = next(s for k, s in net._states.items() if k[4])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", f"{tag}_B{B}_T{T}_{H}x{W}.csv")
st["engine"].profile_csv(path)
st["engine"].profile_csv(path)  # second replay: warm
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in csv.DictReader(open(path)):
    lab = r["label"]
    # strip per-direction / per-level prefixes into shape classes is left to the reader; keep the site key
    k = (r["pass"], lab)
    agg[k][0] += 1
    agg[k][1] += float(r["ms"])
    agg[k][2] += float(r["gflop"])
tot = sum(v[1] for v in agg.values())
print(f"total device time (event-bracketed launches) {tot:.2f} ms for B={B} T={T} {H}x{W}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = v[2] / v[1] if v[1] > 0 else 0.0
    print(f"{v[1]:9.3f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:4d}  {v[1] / v[0] * 1e3:8.1f} us/launch  {tf:7.1f} TF/s  {k[0]} {k[1]}")
