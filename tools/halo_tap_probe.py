"""Per-tap check of the halo-conv engine: weights non-zero for ONE tap only (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
import kernel_cases as K
from refid_b200 import packing
N, H, W, C = 1, 16, 8, 64
x = K.rb(K.g(N, C, H, W, seed=1))
for tap in range(9):
    w = torch.zeros(C, C, 3, 3, device="cuda")
    w[:, :, tap // 3, tap % 3] = K.rb(K.g(C, C, seed=2) / 8)
    ref = F.conv2d(x, w, None, padding=1)
    o = K.run_conv(K.CK_3X3, [K.nhwc(x)], packing.pack_fwd(w).to(torch.bfloat16), C, C, (N, H, W))
    got = K.nchw(o["out"])
    err = (got - ref).abs()
    # which output pixels are wrong
    bad = (err.amax(1)[0] > 0.05)
    print(f"tap {tap} (dy={tap//3-1},dx={tap%3-1}) err {err.max().item():.4f} ref {ref.abs().max().item():.3f} bad pixels {int(bad.sum())}/{H*W}",
          "rows:", sorted(set(bad.nonzero()[:, 0].tolist()))[:16], "cols:", sorted(set(bad.nonzero()[:, 1].tolist())))
