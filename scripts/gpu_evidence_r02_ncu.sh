# Round-2 evidence (part 2): ncu.  Every --set full report is exported to its raw CSV page ON THE BOX and deleted (reports of many kernel
# instances are hundreds of MB; gpurun merges <= 64 MiB).
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_umma_kernel|conv_halo_kernel" -c 400 --csv --log-file gpurun_out/r02_conv_traffic.csv env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_ncu_traffic.log 2>&1; echo "traffic rc=$?"
full() {  # name, kernel regex, extra ncu args..., then the command after --
  name=$1; regex=$2; shift 2
  timeout 1200 ncu --set full --clock-control none -k regex:"$regex" "$@" > gpurun_out/r02_ncu_$name.log 2>&1; rc=$?
  rep=/tmp/r02_$name.ncu-rep
  if [ -f $rep ]; then ncu -i $rep --page raw --csv > gpurun_out/r02_${name}_raw.csv 2>/dev/null; ls -la $rep | cut -c20-80; rm -f $rep; fi
  echo "ncu $name rc=$rc"
}
# the first 44 conv launches of one forward (bg_model: vertical-halo stem, stride-2 convs, the 512->512 trunk, transposed convs, head; then the
# src / tsf stems, SPADE encoders with their merged mlp_shared and (gamma, beta) GEMMs) and the launches from the 88th on (SPADE residual blocks,
# decoders, 128->64 skippers, merged heads); the 9 attention layers; 40 bandwidth kernels; the rasterizer
full conv_umma_a "conv_umma_kernel" -c 44 -o /tmp/r02_conv_umma_a -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16
full conv_umma_b "conv_umma_kernel" --launch-skip 88 -c 44 -o /tmp/r02_conv_umma_b -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16
full attn "conv_halo_kernel|attn_combine" -c 18 -o /tmp/r02_attn -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16
full ops "instnorm_apply|hunfold|hfold|replicate_pad|seg_unfold3" -c 40 -o /tmp/r02_ops -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16
full rast "rasterize_kernel|rast_bin" -c 2 -o /tmp/r02_rast -f python scripts/bench_rasterizer.py 256
gzip -f gpurun_out/r02_launches_bench.csv gpurun_out/r02_conv_traffic.csv
du -sh gpurun_out; ls -la gpurun_out | tail -12
