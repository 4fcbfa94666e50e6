"""A/B of the captured graph with and without parallel branches (GeneratorB200.branch_streams) over batch sizes."""
import json
import sys

import torch

sys.path.insert(0, ".")
from hoig_b200 import synth  # noqa: E402
from hoig_b200.generator import create  # noqa: E402

DIMS = dict(bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12, conv_dim=64, repeat_num=6)   # bench.py CFG
g = create("generator_spade_attn", **DIMS).cuda().eval()
for batch in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 64]:
    inp = {k: v.cuda() for k, v in synth.generator_inputs(batch, seed=1, size=256).items()}
    row = {"batch": batch}
    for mode in ("0", "1"):
        g.branch_streams = mode
        run = g.graphed(inp, with_composite=True)
        for _ in range(3):
            run(**inp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if batch < 32 else 8
        e0.record()
        for _ in range(reps):
            run(**inp)
        e1.record()
        torch.cuda.synchronize()
        row["ms_branches" if mode == "1" else "ms_chain"] = e0.elapsed_time(e1) / reps
        del run
    print(json.dumps(row), flush=True)
