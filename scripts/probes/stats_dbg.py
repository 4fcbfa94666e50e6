import sys, torch
sys.path.insert(0, "/root/repo")
import hoig_b200._lib as L
from tests.test_gpu_ops import _run_conv
case = ("3x3_s1_w256_many_tiles", 3, 256, 64, 64, 3, 1, "conv", dict(stats=True))
for halo in (0, 1):
    for dual in (0, 1):
        L.lib().hoig_set_umma_halo_mode(halo); L.lib().hoig_set_umma_dual_mode(dual)
        for rep in range(2):
            out, ref, st, st_ref = _run_conv(case, torch.bfloat16)
            d = (st.cpu() - st_ref).abs().view(3, 64, 2)
            rel = d / (st_ref.abs().view(3, 64, 2) + 0.5)
            i = rel.argmax()
            print(f"halo {halo} dual {dual} rep {rep}: out maxabs {(out.cpu().float()-ref.float()).abs().max().item():.3e}  stats max abs diff {d.max().item():.4f} worst rel {rel.max().item():.3e} at {i.item()//128, (i.item()//2)%64, i.item()%2} st {st.cpu().view(-1)[i].item():.3f} ref {st_ref.view(-1)[i].item():.3f}")
