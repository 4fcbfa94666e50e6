mkdir -p gpurun_out
for k in 0 1 2 3; do
HOIG_UMMA_2CTA=0 HOIG_UMMA_DEBUG=$k timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_dbg$k.log 2>&1
echo "== pair off, debug $k"; head -n 1 gpurun_out/prof_dbg$k.log; grep -E "k7|Cin128 Cout64 256|convT|k3 s2 Cin64|Cin256 Cout128 128" gpurun_out/prof_dbg$k.log | cut -c1-130
done
