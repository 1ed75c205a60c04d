"""GPU, 2 ranks over NCCL: the data-parallel path.  Each rank runs its own sample; the flat-gradient
all-reduce (mean) inside backward must reproduce the single-process gradient of the concatenated batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch, torch.distributed as dist
import paramgen
from oracle import refid_oracle as O
from refid_b200.arch import FinalBidirectionAttenfusion
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ic, ec, T, H, W = 6, 2, 2, 64, 64
P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
x, ev, gt = paramgen.make_inputs(world, T, H, W, ic, ec)
def run(xs, evs, gts, group):
    torch.manual_seed(100 + rank)  # as the reference's train.py:56: every rank constructs DIFFERENT initial weights
    net = FinalBidirectionAttenfusion(img_chn=ic, ev_chn=ec, num_encoders=3, base_num_channels=32, num_block=1)
    if rank == 0 or group is None:
        net.load_state_dict(P, strict=True)  # only rank 0 holds the intended parameters ...
    net = net.cuda()
    net.grad_sync_group = group  # ... the setter broadcasts them (DDP's constructor behaviour)
    if group is not None:
        for n, p in net.named_parameters():
            assert torch.equal(p.detach().cpu(), P[n]), ("parameters were not broadcast from rank 0", n)
    out = net(x=xs.cuda(), event=evs.cuda())
    torch.sqrt((out - gts.cuda()) ** 2 + 1e-12).mean().backward()
    return {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
mine = run(x[rank:rank + 1], ev[rank:rank + 1], gt[rank:rank + 1], dist.group.WORLD)
full = run(x, ev, gt, None)  # same process, whole batch, no collective
worst = 0.0
for n, g in full.items():
    d = (mine[n] - g).double().norm().item() / max(g.double().norm().item(), 1e-30)
    worst = max(worst, d)
# identical up to bf16 rounding paths that depend on the batch geometry (tile order, atomics)
assert worst < 2e-2, worst
t = torch.tensor([worst], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0: print("MULTI_OK", t.item())
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_flat_gradient_allreduce_matches_full_batch(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
    worst = float(r.stdout.split("MULTI_OK")[1].split()[0])
    print(f"2-rank NCCL flat-gradient all-reduce vs single-process full batch: worst per-parameter rel-L2 difference {worst:.3e}")
