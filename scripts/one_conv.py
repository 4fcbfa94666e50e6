"""Runs ONE conv shape of the generator a few times (for ncu captures / quick timing).
    python scripts/one_conv.py stem|c128_64|convT128|s2_64|heads|res512|spade|mlp128|mlp896 [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hoig_b200 import ops  # noqa: E402
from hoig_b200.packing import pack_conv_weight  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "stem"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dt = torch.float16
B = 64
g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*s):
    return (torch.rand(*s, device="cuda", generator=g) - 0.5).to(dt)


if which == "stem":
    x = rnd(B, 256, 256, 64); w = torch.randn(64, 64, 7, 1, device="cuda") * 0.05
    buf = torch.empty(B, 256, 256, 128, dtype=dt, device="cuda"); out = buf[..., :64]
    kw = dict(kh=7, kw=1, stride=1, pad=3, pad_w=0); tr = False; cout = 64
elif which == "heads":
    x = rnd(B, 256, 256, 128); w = torch.randn(56, 128, 7, 1, device="cuda") * 0.05
    out = torch.empty(B, 256, 256, 56, dtype=dt, device="cuda")
    kw = dict(kh=7, kw=1, stride=1, pad=3, pad_w=0, cout=56); tr = False; cout = None
elif which == "c128_64":
    x = rnd(B, 256, 256, 128); w = torch.randn(64, 128, 3, 3, device="cuda") * 0.05
    out = torch.empty(B, 256, 256, 64, dtype=dt, device="cuda")
    kw = dict(kh=3, kw=3, stride=1, pad=1); tr = False; cout = 64
elif which == "convT128":
    x = rnd(B, 128, 128, 128); w = torch.randn(128, 64, 3, 3, device="cuda") * 0.05
    out = torch.empty(B, 256, 256, 64, dtype=dt, device="cuda")
    kw = dict(kh=3, kw=3, stride=2, pad=1, mode=ops.CONV_TRANSPOSED); tr = True; cout = 64
elif which == "s2_64":
    x = rnd(B, 256, 256, 64); w = torch.randn(128, 64, 3, 3, device="cuda") * 0.05
    out = torch.empty(B, 128, 128, 128, dtype=dt, device="cuda")
    kw = dict(kh=3, kw=3, stride=2, pad=1); tr = False; cout = 128
elif which == "res512":
    x = rnd(B, 32, 32, 512); w = torch.randn(512, 512, 3, 3, device="cuda") * 0.02
    out = torch.empty(B, 32, 32, 512, dtype=dt, device="cuda")
    kw = dict(kh=3, kw=3, stride=1, pad=1, residual=rnd(B, 32, 32, 512)); tr = False; cout = 512
elif which == "spade":
    x = rnd(B, 32, 32, 128); w = torch.randn(1024, 128, 3, 3, device="cuda") * 0.05
    out = torch.empty(B, 32, 32, 512, dtype=dt, device="cuda")
    sx = rnd(B, 32, 32, 512); sst = torch.zeros(B * 512 * 2, dtype=torch.float64, device="cuda"); ops.plane_stats(sx, sst)
    kw = dict(kh=3, kw=3, stride=1, pad=1, cout=1024, spade_x=sx, spade_stats=sst, act=ops.ACT_RELU, bias=torch.zeros(1024, device="cuda")); tr = False; cout = None
elif which == "mlp128":     # merged mlp_shared GEMM over the unfolded segmentation map, 128x128
    x = rnd(B, 128, 128, 64); w = torch.randn(128, 64, 1, 1, device="cuda") * 0.05
    out = torch.empty(B, 128, 128, 128, dtype=dt, device="cuda")
    kw = dict(kh=1, kw=1, stride=1, pad=0, act=ops.ACT_RELU, bias=torch.zeros(128, device="cuda")); tr = False; cout = None
elif which == "mlp896":     # the same at 32x32: 7 SPADE layers in one GEMM
    x = rnd(B, 32, 32, 64); w = torch.randn(896, 64, 1, 1, device="cuda") * 0.05
    out = torch.empty(B, 32, 32, 896, dtype=dt, device="cuda")
    kw = dict(kh=1, kw=1, stride=1, pad=0, act=ops.ACT_RELU, bias=torch.zeros(896, device="cuda")); tr = False; cout = None
else:
    raise SystemExit("unknown shape")
wp = pack_conv_weight(w, dt, transposed=tr)
stats = torch.zeros(B * (cout or 1) * 2, dtype=torch.float64, device="cuda") if cout else None
for _ in range(2):
    ops.conv2d(x, wp, out, stats=stats, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.conv2d(x, wp, out, stats=stats, **kw)
e1.record()
torch.cuda.synchronize()
print(f"{which}: {e0.elapsed_time(e1) / reps:.4f} ms per launch")
