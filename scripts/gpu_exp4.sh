mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 900 $PT tests/test_gpu_umma.py tests/test_gpu_halo.py > gpurun_out/t_umma.log 2>&1; echo "umma+halo rc=$?"; tail -n 2 gpurun_out/t_umma.log
for k in 0 3; do
HOIG_UMMA_DEBUG=$k timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_dbg$k.log 2>&1
echo "== debug $k"; head -n 1 gpurun_out/prof_dbg$k.log; grep -E "conv2d" gpurun_out/prof_dbg$k.log | head -n 22
done
