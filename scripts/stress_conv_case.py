"""Loops ONE conv test case (tests/test_gpu_ops.py::CONV_CASES) many times with fresh random inputs and reports every mismatch against the
CPU emulation -- the bisecting tool for the intermittent failure recorded in DESIGN.md section 9.

    python scripts/stress_conv_case.py 3x3_s2_w256 f16 2000
    HOIG_UMMA_FAST_EPI=0 python scripts/stress_conv_case.py 3x3_s2_w256 f16 2000      # general epilogue
    HOIG_UMMA_MMA_STATS=0 python scripts/stress_conv_case.py 3x3_s2_w256 f16 2000     # shuffle statistics
    compute-sanitizer --tool racecheck python scripts/stress_conv_case.py 3x3_s2_w256 f16 20
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import test_gpu_ops as T  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "3x3_s2_w256"
dtype = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "f16"]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
case = next(c for c in T.CONV_CASES if c[0] == name)
Cout = case[4]
atol = {torch.float16: 3e-3, torch.bfloat16: 2e-2, torch.float32: 2e-4}[dtype]
st_tol = {torch.float16: (1e-3, 0.3), torch.bfloat16: (2e-3, 1.0), torch.float32: (1e-4, 1e-2)}[dtype]
bad = 0
for i in range(reps):
    # _run_conv seeds from hash(name): a different (name, ...) tuple per repetition gives fresh inputs with the same geometry
    c = (f"{name}#{i}",) + case[1:]
    out, ref, st, st_ref = T._run_conv(c, dtype)
    err = (out[..., :Cout].float().cpu() - ref[..., :Cout].float()).abs()
    lim = atol + atol * ref[..., :Cout].float().abs()
    n_bad = int((err > lim).sum())
    st_bad = 0
    if st is not None:
        d = (st.cpu() - st_ref).abs()
        st_bad = int((d > st_tol[1] + st_tol[0] * st_ref.abs()).sum())
    if n_bad or st_bad:
        bad += 1
        idx = torch.nonzero(err > lim)[:8].tolist()
        print(f"rep {i}: {n_bad} output elements and {st_bad} statistics out of tolerance; first bad (n, y, x, c): {idx}; "
              f"max err {err.max().item():.3e}", flush=True)
        if st_bad:
            k = torch.nonzero(d > st_tol[1] + st_tol[0] * st_ref.abs())[:8].flatten().tolist()
            print("   bad statistics entries (index = (n*Cout + c)*2 + {0: sum, 1: sum of squares}):",
                  [(j, float(st.cpu()[j]), float(st_ref[j])) for j in k], flush=True)
print(f"{name} {dtype}: {bad} of {reps} repetitions out of tolerance")
