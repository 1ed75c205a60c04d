"""Single-kernel timing of halo-conv shapes through the C ABI (diagnostic). Usage: python tools/conv_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import kernel_cases as K
from refid_b200 import packing, _lib
import ctypes

def bench(cin, cout, H, W, N, cin2=0, pre=False, reps=20, sv=False):
    x = K.rb(K.g(N, cin + cin2, H, W, seed=1))
    w = K.rb(K.g(cout, cin + cin2, 3, 3, seed=2) / (3 * (cin + cin2) ** 0.5))
    b = K.g(cout, seed=3) * 0.1
    ins = [K.nhwc(x[:, :cin])] + ([K.nhwc(x[:, cin:])] if cin2 else [])
    wp = packing.pack_fwd(w).to(torch.bfloat16)
    r = K.nhwc(K.rb(K.g(N, cout, H, W, seed=4))) if pre else None
    m = K.nhwc(K.rb(K.g(N, cout, H, W, seed=5))) if sv else None  # saved activation: the epilogue applies its LeakyReLU mask
    out = torch.empty(N, H, W, cout, device="cuda", dtype=torch.bfloat16)
    L = _lib.lib()

    def run():
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = L.refid_test_conv(K.CK_3X3, 0, _lib.ptr(ins[0]), cin, _lib.ptr(ins[1] if cin2 else None), cin2, N, H, W, _lib.ptr(wp),
                               ctypes.c_long(wp.shape[0]), wp.shape[1], cout, 0, cout, _lib.ptr(b), _lib.ptr(r), _lib.ptr(m), K.ACT_LRELU,
                               ctypes.c_float(0.1), _lib.ptr(out), None, None, None, None, st)
        _lib.check(rc, "conv")

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    # the C-ABI test entry encodes its tensor maps on every call (tens of microseconds of host time): capture the launches
    # in a CUDA graph so the timing is the device's, not the host's
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    fl = 2.0 * N * H * W * 9 * (cin + cin2) * cout
    if hasattr(L, "refid_debug_halo_timing"):
        buf = (ctypes.c_longlong * (148 * 8))()
        L.refid_debug_halo_timing(buf)
        rows = [buf[i * 8:i * 8 + 6] for i in range(148)]
        worst = max(rows, key=lambda r: r[0])
        print("   MMA warp cycles (slowest CTA): total %d, acc_empty wait %d, A wait %d, B wait %d, issue blocks %d, items %d" % tuple(worst))
    print(f"cin {cin}+{cin2} cout {cout} {N}x{H}x{W} pre={pre} sv={sv}: {us:8.1f} us  {fl / us / 1e6:7.1f} TF/s  (device time, CUDA-graph replay)", flush=True)

print("REFID_EPI_L2PF =", os.environ.get("REFID_EPI_L2PF"))
if os.environ.get("CONV_BENCH_ONLY") == "enc0":
    bench(32, 128, 256, 256, 8)     # enc0_in: both directions' level-0 in-convs stacked (one step's worth of the all-T launch)
    bench(32, 128, 256, 256, 32)
    bench(32, 64, 256, 256, 8)
    bench(64, 128, 256, 256, 8)
    sys.exit(0)
bench(64, 64, 256, 256, 8)
bench(64, 64, 256, 256, 8, pre=True)
bench(64, 64, 256, 256, 8, pre=True, sv=True)
bench(128, 128, 128, 128, 8, pre=True)
bench(128, 128, 128, 128, 8, pre=True, sv=True)
bench(64, 64, 256, 256, 8, cin2=64)
bench(128, 128, 128, 128, 8)
bench(128, 128, 128, 128, 8, cin2=128)
bench(256, 256, 64, 64, 8)
bench(256, 256, 64, 64, 8, cin2=256)
bench(32, 32, 256, 256, 8)
