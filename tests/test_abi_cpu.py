"""No-GPU checks of the boundary: the library loads, exports every symbol the header declares,
argument validation fails loudly, and the product never reaches for the oracle."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hoig_b200 import _lib, build
    build.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "hoig_b200.h")).read()
    declared = set(re.findall(r"\b(hoig_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 24
    from hoig_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.hoig_version()


def test_sass_contains_blackwell_instructions():
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "hoig_b200", "_C", "libhoig_b200.so")],
                         capture_output=True, text=True).stdout
    if not out:
        pytest.skip("cuobjdump unavailable")
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "LDGSTS"):
        assert mnemonic in out, mnemonic
    # The convolutions run on tcgen05 only.  Legacy warp-level HMMA may appear in exactly two places: inside the tcgen05 conv kernel, where
    # it reduces the epilogue's per-plane statistics (umma_common.cuh colsum16), and in attn_combine_tc_kernel, whose per-tile window GEMMs
    # (64 x <=256 x 64 per slab, A built on the fly in shared memory) are warp-level by design.
    for fn in out.split("Function :")[1:]:
        body = fn.replace("UTCHMMA", "")
        n_legacy = body.count("HMMA.")
        if n_legacy:
            head = fn.split("\n", 1)[0]
            assert ("UTCHMMA" in fn and "conv_umma_kernel" in head) or "attn_combine_tc_kernel" in head, head
            assert n_legacy <= 64, (head, n_legacy)


def test_argument_validation_without_gpu(lib):
    from hoig_b200._lib import ConvDesc
    d = ConvDesc()
    assert lib.hoig_conv2d(ctypes.byref(d), None) == -1
    assert b"null pointer" in lib.hoig_last_error()
    assert lib.hoig_conv2d(None, None) == -1
    assert lib.hoig_rasterize_fim_wim(None, 1, 1, 64, 0.1, 100.0, 1, None, None, None, None, 0, None) == -1
    # hidden, ldh, Chid, w2, b2, src, lds, flow, tgt, ldt, dst, ldd, dtype, N, h, C, k, stream
    assert lib.hoig_attn_finish(None, 0, 128, None, None, None, 0, None, None, 0, None, 0, 1, 1, 8, 8, 5, None, 0, None) == -1
    r, c = ctypes.c_int(), ctypes.c_int()
    lib.hoig_conv_packed_dims(0, 3, 7, 7, 64, 1, 3, ctypes.byref(r), ctypes.byref(c))
    assert (r.value, c.value) == (16, 3136)
    lib.hoig_conv_packed_dims(2, 128, 5, 5, 1024, 5, 0, ctypes.byref(r), ctypes.byref(c))
    assert (r.value, c.value) == (128, 25600)
    lib.hoig_conv_packed_dims(1, 64, 3, 3, 128, 2, 1, ctypes.byref(r), ctypes.byref(c))     # 4 parity blocks x 4 input taps
    assert (r.value, c.value) == (256, 512)
    lib.hoig_conv_packed_dims(1, 16, 3, 3, 8, 2, 1, ctypes.byref(r), ctypes.byref(c))
    assert (r.value, c.value) == (64, 64)


def test_ops_fail_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hoig_b200 import ops
    with pytest.raises((RuntimeError, ValueError)):
        ops.rasterize(torch.zeros(1, 1, 3, 3), 64)
    from hoig_b200.generator import create
    from hoig_b200 import synth
    g = create("generator_base", bg_dim=8, img_dim=3, obj_dim=3, conv_dim=16)
    with pytest.raises((RuntimeError, ValueError)):
        g(**synth.generator_inputs(1, size=64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "hoig_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "hoig_oracle" not in txt, f
    code = "import sys; import hoig_b200, hoig_b200.generator, hoig_b200.ops, hoig_b200.renderer, hoig_b200.compat; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_packing_layout():
    from hoig_b200.packing import pack_conv_weight
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = pack_conv_weight(w, torch.float32)
    assert p.shape == (16, 128)
    # k = (r*KW + s)*Cin_pad + c with Cin_pad = 8
    assert p[1, (1 * 3 + 2) * 8 + 2] == w[1, 2, 1, 2]
    assert p[0, 3] == 0 and p[2:].abs().sum() == 0
    wt = torch.arange(3 * 2 * 3 * 3, dtype=torch.float32).reshape(3, 2, 3, 3)   # ConvTranspose2d (Cin,Cout,kh,kw)
    pt = pack_conv_weight(wt, torch.float32, transposed=True)
    assert pt.shape == (16, 64)                            # 4 parities x Cout=2 rows (pad 16), 4 taps x Cin_pad=8 cols (pad 64)
    assert pt[0 * 2 + 1, 0 * 8 + 2] == wt[2, 1, 1, 1]      # parity (0,0) uses only tap (0,0) with the centre weight
    assert pt[0 * 2 + 1, 8:].abs().sum() == 0
    from hoig_b200.packing import parity_block
    assert [parity_block(a, b) for a in (0, 1) for b in (0, 1)] == [0, 1, 3, 2]     # row blocks ordered (0,0),(0,1),(1,1),(1,0)
    assert pt[2 * 2 + 1, (1 * 2 + 0) * 8 + 2] == wt[2, 1, 0, 2]   # parity (1,1) = block 2, tap (dy,dx)=(1,0): kernel index (0,2)
    assert pt[3 * 2 + 1, (0 * 2 + 1) * 8:(0 * 2 + 2) * 8].abs().sum() == 0        # parity (1,0) = block 3 never uses a dx = 1 tap
    from tests.emu_ops import unpack_transposed
    assert torch.equal(unpack_transposed(pt, 2, 3, 3, 8, 1)[:, :3], wt.permute(1, 0, 2, 3))
