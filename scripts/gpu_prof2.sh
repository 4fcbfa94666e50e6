mkdir -p gpurun_out
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:instnorm_apply -s 300 -c 6 -o gpurun_out/instnorm_full python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_in.log 2>&1; echo "ncu rc=$?"
grep -E "instnorm|forward" gpurun_out/prof_convs_b64.log
