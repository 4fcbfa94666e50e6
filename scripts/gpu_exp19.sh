mkdir -p gpurun_out
for k in 0 1; do
HOIG_UMMA_DEBUG=$k timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_dbg$k.log 2>&1
echo "== debug $k"; head -n 24 gpurun_out/prof_dbg$k.log | cut -c1-125
done
