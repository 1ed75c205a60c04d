"""Run every single-kernel parity case on the GPU and print a table (diagnostic; not a test)."""
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import kernel_cases  # noqa: E402

res = {}
only = sys.argv[1:]
for name, fn in kernel_cases.CASES.items():
    if only and not any(o in name for o in only):
        continue
    try:
        err, ref = fn()
        ok = err == err and err <= 1.5e-2 * max(ref, 1e-6)
        res[name] = {"err": err, "ref_max": ref, "ok": bool(ok)}
    except Exception as e:  # noqa: BLE001
        res[name] = {"error": repr(e)[:300]}
        traceback.print_exc()
    print(name, res[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "kernel_probe.json"), "w"), indent=1)
print("PASS" if all(r.get("ok") for r in res.values()) else "FAIL")
