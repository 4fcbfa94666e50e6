mkdir -p gpurun_out
timeout 600 python scripts/bench_conditions.py 64 > gpurun_out/bench_conditions.log 2>&1; echo "cond rc=$?"; tail -n 3 gpurun_out/bench_conditions.log | cut -c1-1500
