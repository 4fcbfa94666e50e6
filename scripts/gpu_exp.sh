mkdir -p gpurun_out
for k in 0 1 2 4 8; do HOIG_UMMA_PREFETCH_TILES=$k timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_pf$k.log 2>&1; echo "prefetch=$k: $(head -n 1 gpurun_out/prof_pf$k.log)"; grep -E "Cin128 Cout64 256x256|k7 s1 Cin64 Cout64|convT k3 s2 Cin128|Cin512 Cout512" gpurun_out/prof_pf$k.log | cut -c1-120; done
