# round-2 run 1: all gpu tests, bench f16 vs bf16 back to back, per-shape profile
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 1800 $PT tests > gpurun_out/r2_t_all.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/r2_t_all.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_f16.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_f16.log | cut -c1-800
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dtype bf16 > gpurun_out/r2_bench_bf16.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_bf16.log | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_f16b.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_f16b.log | cut -c1-400
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 1 > gpurun_out/r2_bench_b1.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/r2_bench_b1.log | cut -c1-400
timeout 600 python scripts/profile_convs.py 64 f16 > gpurun_out/r2_prof_convs_b64.log 2>&1; head -n 45 gpurun_out/r2_prof_convs_b64.log | cut -c1-150
