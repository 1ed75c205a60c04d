"""Host time to ENQUEUE one training step (forward + loss + backward, no synchronisation) against its device time
(diagnostic: how far the launching thread runs ahead of the GPU).  Usage: python tools/host_launch_probe.py [B T H W]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refid_b200.arch import FinalBidirectionAttenfusion
from refid_b200.losses import CharbonnierLoss

B, T, H, W = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (8, 23, 256, 256)
net = FinalBidirectionAttenfusion(img_chn=26, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1, num_residual_blocks=2).cuda()
x = torch.rand(B, 26, H, W, device="cuda")
ev = torch.randn(B, T, 2, H, W, device="cuda")
gt = torch.rand(B, T, 3, H, W, device="cuda")
cri = CharbonnierLoss()


def step():
    for p in net.parameters():
        p.grad = None
    cri(net(x=x, event=ev), gt).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
host, dev = [], []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    dev.append(e0.elapsed_time(e1))
st = next(s for k, s in net._states.items() if k[4])
print("cuda graphs:", st["engine"].graph_stats())
print(f"B={B} T={T} {H}x{W}: host enqueue {sorted(host)[2]:.1f} ms per step, device {sorted(dev)[2]:.1f} ms per step "
      f"({os.cpu_count()} host CPUs)")
