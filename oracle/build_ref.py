"""Builds the REFERENCE's own CUDA kernels into oracle/_ref -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference, nvcc and torch headers):
    python oracle/build_ref.py

The three reference extensions (thirdparty/block_extractor,
thirdparty/local_attn_reshape, thirdparty/neural_renderer/.../cuda/rasterize*)
do not compile against torch 2.x as shipped: every
``AT_DISPATCH_FLOATING_TYPES(x.type(), ...)`` needs ``x.scalar_type()``
(10 sites).  This script copies the sources to a scratch
dir under /tmp, applies that mechanical substitution (device code untouched),
compiles them for sm_100a with the same nvcc as the product, and leaves ONLY
the resulting .so files in oracle/_ref/ (git-ignored; they travel to the GPU
box with the snapshot).  No reference source enters the repository.

tests/ load these modules (when present) as the same-toolchain GPU oracle:
bit-exact fim, and BlockExtractor / LocalAttnReshape forward values.
"""
import os
import re
import shutil
import sys
import tempfile

REF = "/root/reference/HOIG_HOv3/thirdparty"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

EXTS = {
    "ref_block_extractor_cuda": ("block_extractor", ["block_extractor_cuda.cc", "block_extractor_kernel.cu"],
                                 ["block_extractor_kernel.cuh"]),
    "ref_local_attn_reshape_cuda": ("local_attn_reshape", ["local_attn_reshape_cuda.cc", "local_attn_reshape_kernel.cu"],
                                    ["local_attn_reshape_kernel.cuh"]),
    "ref_rasterize_cuda": ("neural_renderer/neural_renderer/cuda", ["rasterize_cuda.cpp", "rasterize_cuda_kernel.cu"], []),
}


def main():
    if not os.path.isdir(REF):
        print("reference not present; nothing to build")
        return 0
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    os.makedirs(OUT, exist_ok=True)
    for name, (sub, srcs, hdrs) in EXTS.items():
        tmp = tempfile.mkdtemp(prefix="hoig_ref_")
        paths = []
        for f in srcs + hdrs:
            txt = open(os.path.join(REF, sub, f)).read()
            txt = re.sub(r"(AT_DISPATCH_FLOATING_TYPES\(\s*\w+)\.type\(\)", r"\1.scalar_type()", txt)
            txt = txt.replace(".data<", ".data_ptr<")
            dst = os.path.join(tmp, f)
            open(dst, "w").write(txt)
            if f in srcs:
                paths.append(dst)
        bdir = os.path.join(tmp, "build")
        os.makedirs(bdir)
        load(name=name, sources=paths, build_directory=bdir, verbose=False,
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"], is_python_module=False)
        shutil.copy(os.path.join(bdir, name + ".so"), os.path.join(OUT, name + ".so"))
        shutil.rmtree(tmp, ignore_errors=True)
        print("built", name)
    return 0


if __name__ == "__main__":
    sys.exit(main())
