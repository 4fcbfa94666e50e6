mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -s -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 600 $PT tests/test_gpu_halo.py -k "attn_combine or commuted" > gpurun_out/t_halo.log 2>&1; echo "halo rc=$?"; tail -n 2 gpurun_out/t_halo.log
for k in 0 600 1100 2100 5000; do
HOIG_STATS_EPILOGUE_MIN_K=$k timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_k$k.log 2>&1
echo "== min_k $k"; head -n 1 gpurun_out/prof_k$k.log; grep -E "convT|plane_stats|attn_combine|k3 s2" gpurun_out/prof_k$k.log
done
