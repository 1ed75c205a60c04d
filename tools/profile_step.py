"""One warm-up + one measured training step of the hot path at a reduced batch (same per-tile shapes as the bench
workload) for ncu launch lists / full captures.  Usage: python tools/profile_step.py [B] [T] [H] [W]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from refid_b200.arch import FinalBidirectionAttenfusion  # noqa: E402

B, T, H, W = [int(v) for v in (sys.argv[1:5] + ["2", "2", "256", "256"][len(sys.argv) - 1:])]
torch.manual_seed(0)
net = FinalBidirectionAttenfusion(img_chn=26, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1)
bench.init_params(net)
net = net.cuda()
x, ev, gt = bench.make_inputs(B, T, H, W, 26, 2, seed=1234, device="cuda")
for it in range(2):
    for p in net.parameters():
        p.grad = None
    out = net(x=x, event=ev)
    loss = torch.sqrt((out - gt) ** 2 + 1e-12).mean()
    loss.backward()
    torch.cuda.synchronize()
print("loss", loss.item())
