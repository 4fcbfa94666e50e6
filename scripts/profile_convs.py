"""Per-shape timing of every conv launch in one generator forward (CUDA events on the launching stream)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CFG, GFLOP_PER_IMAGE  # noqa: E402
from hoig_b200 import _lib, synth  # noqa: E402
from hoig_b200.generator import create  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dtype = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "f16"]
g = create("generator_spade_attn", dtype=dtype, **CFG).cuda().eval()
inp = {k: v.cuda() for k, v in synth.generator_inputs(B, seed=1, size=256).items()}
for _ in range(0 if os.environ.get("HOIG_PROFILE_SINGLE") else 2):      # HOIG_PROFILE_SINGLE=1: exactly one forward (ncu captures)
    g(**inp)
torch.cuda.synchronize()
_lib.recorder.reset(timing=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g(**inp); e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1)
rows = sorted(_lib.recorder.by_tag().items(), key=lambda kv: -kv[1][1])
print(f"forward B={B} {dtype}: {total:.2f} ms  ({B / total * 1e3:.1f} img/s), {_lib.recorder.launches} launches")
acc = 0.0
for (name, tag), (n, ms) in rows[:60]:
    acc += ms
    extra = ""
    if tag and name == "hoig_instnorm_apply":
        import re
        m = re.match(r"C(\d+) (\d+)x(\d+) N(\d+) gb=(\d) res=(\d)", tag)
        c, h, w, nn, gbf, rs = map(int, m.groups())
        by = nn * h * w * c * 2 * (2 + 2 * gbf + rs)
        extra = f" {by * n / ms / 1e6:8.1f} GB/s"
    elif tag:
        import re
        m = re.match(r"(\w+) k(\d+)x(\d+) s(\d+) Cin(\d+) Cout(\d+) (\d+)x(\d+)->(\d+)x(\d+) N(\d+)", tag)
        mode, kh, kw, s, cin, cout, h, w, oh, ow, nn = m.group(1), *map(int, m.groups()[1:])
        # FLOPs of the GEMM as the layer defines it (2 * MACs over the operand shapes the kernel sees: input channels padded to 8 / the
        # 7-tap unfold padded to 64; a transposed conv counts its 9 live taps per input pixel, not the 16 blocks of the packed matrix)
        fl = 2.0 * nn * (h * w if mode == "convT" else oh * ow) * cout * cin * kh * kw
        extra = f" {fl * n / ms / 1e9:8.1f} TFLOP/s"
    print(f"{ms:9.3f} ms {100 * ms / total:5.1f}% n={n:3d} {name.replace('hoig_', ''):16s} {tag or ''}{extra}")
print(f"listed {acc:.1f} ms of {total:.1f}")
