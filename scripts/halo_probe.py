"""Probe: hoig_conv2d_halo A-operand variants (shifted descriptors vs one box per tap): error vs torch and timing."""
import sys
import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from hoig_b200 import _lib, ops
from hoig_b200.packing import pack_conv_weight


def run(variant, n, hp, c, k, cout, dtype, check=True, two=False):
    _lib.lib().hoig_set_halo_variant(variant)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, hp, hp, c, generator=g).to(dtype).cuda()
    w = (torch.randn(cout, c, k, k, generator=g) * 0.05)
    wp = pack_conv_weight(w, dtype).cuda()
    out = torch.zeros(n, hp, hp, cout, dtype=dtype, device="cuda")
    segs = [(x, wp, out)]
    if two:
        x2 = torch.randn(n, hp + 4, hp + 4, c, generator=g).to(dtype).cuda()
        out2 = torch.zeros(n, hp + 4, hp + 4, cout, dtype=dtype, device="cuda")
        segs.append((x2, wp, out2))
    ops.conv2d_halo(segs, k, k, cout)
    torch.cuda.synchronize()
    res = {}
    if check:
        for i, (xx, _, oo) in enumerate(segs):
            ref = F.conv2d(xx.float().permute(0, 3, 1, 2), w.to(dtype).float().cuda(), None, padding=k // 2).permute(0, 2, 3, 1)
            r = k // 2
            d = (oo.float() - ref)[:, r:-r, r:-r]
            res[f"seg{i}_maxabs"] = d.abs().max().item()
            res[f"seg{i}_rel"] = (d.norm() / ref[:, r:-r, r:-r].norm()).item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        ops.conv2d_halo(segs, k, k, cout)
    e0.record()
    for _ in range(5):
        ops.conv2d_halo(segs, k, k, cout)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = sum(2.0 * xx.shape[0] * xx.shape[1] * xx.shape[2] * c * k * k * cout for xx, _, _ in segs)
    res["ms"] = ms
    res["TFLOPs"] = flops / ms / 1e9
    return res


if __name__ == "__main__":
    for variant in (2, 0, 1):
        for (n, hp, c, k) in ((2, 20, 64, 5), (2, 36, 128, 3)):
            try:
                print("variant", variant, (n, hp, c, k), run(variant, n, hp, c, k, 128, torch.bfloat16), flush=True)
            except Exception as e:
                print("variant", variant, "FAILED", repr(e)[:300], flush=True)
                sys.exit(1 if variant == 2 else 0)
    for variant in (2, 0):
        for (n, hp, c) in ((64, 36, 512), (64, 132, 128), (64, 68, 256)):
            print("timing variant", variant, (n, hp, c), run(variant, n, hp, c, 5, 128, torch.bfloat16, check=False, two=True), flush=True)
