"""Run golden cases of the whole network on the GPU; compare output, loss, parameter gradients with the golden vectors
and named intermediates with the CPU oracle (diagnostic; not a test)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import golden_util  # noqa: E402
import paramgen  # noqa: E402
from oracle import refid_oracle as O  # noqa: E402
from oracle import refid_oracle_bf16 as OB  # noqa: E402
from refid_b200 import _lib  # noqa: E402
from refid_b200.arch import FinalBidirectionAttenfusion  # noqa: E402


def run(case, intermediates=True, grads=True):
    B, T, H, W, ic, ec, x5d = golden_util.CASES[case]
    gold = golden_util.load(case)
    P = paramgen.make_params(O.param_shapes(ic, ec), seed=0)
    x, ev, gt = paramgen.make_inputs(B, T, H, W, ic, ec, x5d=x5d)
    net = FinalBidirectionAttenfusion(img_chn=ic, ev_chn=ec, num_encoders=3, base_num_channels=32, num_block=1,
                                      num_residual_blocks=2)
    net.load_state_dict(P, strict=True)
    net = net.cuda()
    out = net(x=x.cuda(), event=ev.cuda())
    torch.cuda.synchronize()
    flag = _lib.abort_flag()
    res = {"abort_flag": flag}
    res["out_err"] = (out.detach().cpu() - gold["out"]).abs().max().item()
    res["out_max"] = gold["out"].abs().max().item()
    print(case, "out err", res["out_err"], "of", res["out_max"], "abort", hex(flag), flush=True)
    if intermediates:
        rec = {}
        with torch.no_grad():
            O.forward(P, x, ev, rec)
        st = next(iter(net._states.values()))
        eng = st["engine"]
        names = sorted(rec, key=lambda n: (n.split(".")[0] != "head_img", n))
        worst = []
        for n in rec:
            try:
                mine = eng.debug_tensor(n).cpu()
            except RuntimeError:
                continue
            ref = rec[n]
            e = (mine - ref).abs().max().item()
            worst.append((n, e, ref.abs().max().item()))
        for n, e, m in worst:
            flagc = "  <<<" if e > 0.03 * max(m, 1e-3) else ""
            print(f"   {n:24s} err {e:.4g} max {m:.4g}{flagc}")
        res["intermediates"] = {n: [e, m] for n, e, m in worst}
    if grads:
        loss = torch.sqrt((out - gt.cuda()) ** 2 + 1e-12).mean()
        loss.backward()
        torch.cuda.synchronize()
        res["abort_flag_bwd"] = _lib.abort_flag()
        res["loss_err"] = abs(loss.item() - gold["loss"])
        g = {}
        for n, p in net.named_parameters():
            g[n] = p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(p).cpu()
        rows = []
        for n in gold["names"]:
            if n in gold["dead"]:
                continue
            gn = gold["grad_norm"][n]
            mine = g[n].double().norm().item()
            idx = paramgen.grad_sample_index(n, g[n].numel())
            scale = gn / max(g[n].numel(), 1) ** 0.5
            serr = (g[n].flatten()[idx] - gold["grad_samples"][n]).abs().max().item()
            rows.append((n, mine, gn, abs(mine - gn) / max(gn, 1e-12), serr / max(scale, 1e-12)))
        # fixed-cotangent VJP: removes the sign(pred-gt) discontinuity of the Charbonnier gradient from the comparison
        cot = (torch.randn(gold["out"].shape, generator=torch.Generator().manual_seed(7)) / gold["out"].numel())
        cot = cot.bfloat16().float()
        ob_out, ob_g = OB.vjp(P, x, ev, cot)
        net.zero_grad(set_to_none=True)
        out2 = net(x=x.cuda(), event=ev.cuda())
        (out2 * cot.cuda()).sum().backward()
        torch.cuda.synchronize()
        res["out_err_vs_bf16_oracle"] = (out2.detach().cpu() - ob_out).abs().max().item()
        print("   out err vs bf16-emulating oracle", res["out_err_vs_bf16_oracle"])
        vj = []
        for n in gold["names"]:
            if n in gold["dead"]:
                continue
            ref = ob_g[n].double()
            mine = dict(net.named_parameters())[n].grad.detach().cpu().double()
            vj.append((n, ((mine - ref).norm() / ref.norm().clamp_min(1e-30)).item(), ref.norm().item()))
        vj.sort(key=lambda r: -r[1])
        for n, e, m in vj[:12]:
            print(f"   VJP relL2 {n:64s} {e:.4g}  (|g| {m:.3g})")
        res["vjp_rel_l2_max"] = vj[0][1]
        res["vjp_rel_l2"] = {n: e for n, e, m in vj}
        _, _, og = O.loss_and_grads(P, x, ev, gt)
        l2 = []
        for n in gold["names"]:
            if n in gold["dead"]:
                continue
            ref = og[n].double()
            l2.append((n, ((g[n].double() - ref).norm() / ref.norm().clamp_min(1e-30)).item(), ref.norm().item()))
        l2.sort(key=lambda r: -r[1])
        for n, e, m in l2[:4]:
            print(f"   relL2 {n:64s} {e:.4g}  (|g| {m:.3g})")
        res["grad_rel_l2_max"] = l2[0][1]
        res["grad_rel_l2"] = {n: e for n, e, m in l2}
        rows.sort(key=lambda r: -max(r[3], r[4] / 10))
        print("   loss err", res["loss_err"], "abort", hex(res["abort_flag_bwd"]))
        for r in rows[:8]:
            print(f"   grad {r[0]:64s} norm {r[1]:.4g} vs {r[2]:.4g} rel {r[3]:.3g} sample/rms {r[4]:.3g}")
        res["grad_worst"] = [[r[0], r[3], r[4]] for r in rows[:25]]
        res["grad_norm_rel_max"] = max(r[3] for r in rows)
        res["grad_sample_rel_max"] = max(r[4] for r in rows)
        print("   grad norm rel max", res["grad_norm_rel_max"], "sample/rms max", res["grad_sample_rel_max"], flush=True)
    return res


if __name__ == "__main__":
    cases = sys.argv[1:] or list(golden_util.CASES)
    allres = {}
    for c in cases:
        try:
            allres[c] = run(c)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            allres[c] = {"error": repr(e)[:500]}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(allres, open(os.path.join(ROOT, "gpurun_out", "net_probe.json"), "w"), indent=1)
