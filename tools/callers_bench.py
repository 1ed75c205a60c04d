"""Device times of the two SURVEY 8f rows built so far against the torch calls the reference makes (diagnostic):
Charbonnier loss fwd+bwd on the (8,23,3,256,256) output, and clip_grad_norm_(0.01) + AdamW.step on the network's 183
parameter tensors.  Usage: python tools/callers_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from refid_b200 import losses, optim
from refid_b200.arch import FinalBidirectionAttenfusion


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


shape = (8, 23, 3, 256, 256)
pred = torch.rand(shape, device="cuda", requires_grad=True)
gt = torch.rand(shape, device="cuda")
cri = losses.CharbonnierLoss()


def loss_fused():
    pred.grad = None
    cri(pred, gt).backward()


def loss_torch():
    pred.grad = None
    torch.sqrt((pred - gt) ** 2 + 1e-12).mean().backward()


n = pred.numel()
t_f, t_t = timeit(loss_fused), timeit(loss_torch)
print(f"Charbonnier fwd+bwd, {n} elements: fused {t_f:.1f} us ({12.0 * n / t_f / 1e6:.2f} TB/s of its 12 B/element), torch ops {t_t:.1f} us")

net = FinalBidirectionAttenfusion(img_chn=26, ev_chn=2, num_encoders=3, base_num_channels=32, num_block=1, num_residual_blocks=2).cuda()
params = [p for p in net.parameters()]
for p in params:
    p.grad = torch.randn_like(p) * 1e-3
ne = sum(p.numel() for p in params)
o_f = optim.ClipAdamW(params, lr=2e-4, betas=(0.9, 0.99), weight_decay=1e-4)
o_t = torch.optim.AdamW(params, lr=2e-4, betas=(0.9, 0.99), weight_decay=1e-4)
o_tf = torch.optim.AdamW(params, lr=2e-4, betas=(0.9, 0.99), weight_decay=1e-4, fused=True)


def step_fused():
    o_f.clip_grad_norm_(0.01)
    o_f.step()


def step_torch(o):
    def f():
        torch.nn.utils.clip_grad_norm_(params, 0.01)
        o.step()
    return f


t_f, t_t, t_tf = timeit(step_fused), timeit(step_torch(o_t)), timeit(step_torch(o_tf))
print(f"clip 0.01 + AdamW step, {len(params)} tensors / {ne} elements: fused {t_f:.1f} us ({32.0 * ne / t_f / 1e6:.2f} TB/s of its "
      f"32 B/element), torch clip_grad_norm_ + AdamW(foreach) {t_t:.1f} us, + AdamW(fused=True) {t_tf:.1f} us  (wall-clock per call incl. host)")
