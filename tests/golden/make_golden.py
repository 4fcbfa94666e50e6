"""Generates tests/golden/*.npz by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

What runs: the reference ``Generator`` class exactly as shipped
(/root/reference/HOIG_HOv3/models/networks/generator.py) with
 * import stubs for h5py / smplx (pulled in by models/networks/__init__.py:1),
 * the two CUDA-only ops (BlockExtractor, LocalAttnReshape -- their Python
   wrappers raise NotImplementedError on CPU tensors, block_extractor.py:23)
   served by the oracle's C restatement through the reference's own pybind
   call signature ``forward(source, flow, output, k)``,
 * ``Tensor.cuda`` neutralised (generator.py:487 calls it unconditionally).
Weights come from ``oracle.generator_ref.init_state_dict`` and are loaded with
``load_state_dict(strict=True)``, which also pins the 425-entry key layout.

Outputs (small, committed): per config a strided sample of each of the 10
generator outputs, full-tensor sums, and the state_dict key/shape listing.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/HOIG_HOv3"


def import_reference():
    import oracle

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return object

    for m in ["h5py", "smplx", "smplx.lbs", "smplx.utils", "smplx.vertex_ids", "smplx.vertex_joint_selector"]:
        sys.modules.setdefault(m, _Any(m))

    be = types.ModuleType("block_extractor_cuda")

    def be_forward(source, flow, output, k):
        output.copy_(torch.from_numpy(oracle.block_extract(source.numpy(), flow.numpy(), k)))
        return 1

    be.forward = be_forward
    la = types.ModuleType("local_attn_reshape_cuda")

    def la_forward(inputs, output, k):
        output.copy_(torch.from_numpy(oracle.local_attn_reshape(inputs.numpy(), k)))
        return 1

    la.forward = la_forward
    sys.modules["block_extractor_cuda"] = be
    sys.modules["local_attn_reshape_cuda"] = la
    sys.path.insert(0, REF)
    from models.networks.generator import Generator  # noqa
    import thirdparty.block_extractor.block_extractor as bem
    import thirdparty.local_attn_reshape.local_attn_reshape as lam

    # the reference wrappers refuse CPU tensors before reaching the op module;
    # lie about is_cuda by routing through thin Function subclasses
    class _BE(torch.nn.Module):
        def __init__(self, k):
            super().__init__()
            self.k = k

        def forward(self, source, flow):
            out = flow.new_zeros(source.shape[0], source.shape[1], self.k * flow.shape[2], self.k * flow.shape[3])
            be_forward(source.contiguous(), flow.contiguous(), out, self.k)
            return out

    class _LA(torch.nn.Module):
        def forward(self, x, k):
            out = x.new_zeros(x.shape[0], 1, k * x.shape[2], k * x.shape[3])
            la_forward(x.contiguous(), out, k)
            return out

    torch.Tensor.cuda = lambda self, *a, **k: self
    return Generator, _BE, _LA


CONFIGS = {
    # name: (ctor kwargs, image size, batch, stride of the committed sample)
    "small": (dict(bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12, conv_dim=16, repeat_num=6), 64, 2, 4),
    "full": (dict(bg_dim=8, img_dim=3, obj_dim=3, img_cond_dim=3, obj_cond_dim=12, conv_dim=64, repeat_num=6), 256, 1, 8),
}


def main():
    from oracle import generator_ref as gr
    from hoig_b200 import synth

    Generator, _BE, _LA = import_reference()
    for name, (kw, size, B, stride) in CONFIGS.items():
        torch.manual_seed(0)
        g = Generator(**kw, spade_layers=[1, 1, 0, 0], attn_layers=list(range(1, 10)))
        for L in range(1, 10):
            a = getattr(g, "attn_%d" % L)
            a.extractor = _BE(a.kernel_size)
            a.reshape = _LA()
        sd = gr.init_state_dict(seed=0, jitter=0.05, **kw)
        ref_keys = [(k, tuple(v.shape)) for k, v in g.state_dict().items()]
        assert ref_keys == [(k, tuple(s)) for k, s in gr.state_dict_spec(**kw)], "state_dict layout mismatch"
        g.load_state_dict(sd, strict=True)
        g.eval()
        inp = synth.generator_inputs(B, seed=1, size=size)
        with torch.no_grad():
            outs = g(**inp)
        rec = {}
        for i, o in enumerate(outs):
            o = o.double()
            rec[f"out{i}_sample"] = o[:, :, ::stride, ::stride].float().numpy()
            rec[f"out{i}_sum"] = np.array([o.sum().item(), o.abs().sum().item()])
        np.savez_compressed(os.path.join(HERE, f"generator_{name}.npz"), **rec)
        if name == "full":
            with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
                json.dump([[k, list(s)] for k, s in ref_keys], f)
        print(name, "ok", [tuple(o.shape) for o in outs][:3], "...")


if __name__ == "__main__":
    main()
