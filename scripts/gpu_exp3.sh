mkdir -p gpurun_out
for k in 0 1 2 3; do
HOIG_UMMA_DEBUG=$k timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_dbg$k.log 2>&1
echo "== debug $k"; head -n 1 gpurun_out/prof_dbg$k.log; grep -E "conv2d " gpurun_out/prof_dbg$k.log | head -n 24
done
