mkdir -p gpurun_out
for r in 1 2; do for c in 1 0; do
HOIG_UMMA_CONTIG=$c timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_contig${c}_$r.log 2>&1
echo "== contig $c run $r"; head -n 1 gpurun_out/prof_contig${c}_$r.log; grep -E "conv2d " gpurun_out/prof_contig${c}_$r.log | head -n 12 | cut -c1-120
done; done
