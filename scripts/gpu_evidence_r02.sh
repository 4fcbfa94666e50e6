# Round-2 evidence (part 1): tests, smoke, bench arms, per-shape profile.  Small logs only (gpurun merges <= 64 MiB).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 1800 $PT tests > gpurun_out/r02_t_all.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r02_t_all.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.log 2>&1; echo "bench rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 900 python bench.py --mode train --steps 3 --warmup 1 > gpurun_out/r02_bench_train.log 2>&1; echo "bench train rc=$?"
timeout 300 python bench.py --mode train --batch 16 --steps 2 --warmup 1 > gpurun_out/r02_bench_train_b16.log 2>&1; echo "bench train b16 rc=$?"
timeout 200 python scripts/branch_ab.py 1 4 64 > gpurun_out/r02_branch_ab.log 2>&1
timeout 600 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_prof_convs_b64.log 2>&1
timeout 600 python scripts/bench_conditions.py 64 > gpurun_out/r02_conditions.log 2>&1
tail -n 1 gpurun_out/r02_bench.log | cut -c1-600; tail -n 1 gpurun_out/r02_bench_ref.log | cut -c1-300; tail -n 1 gpurun_out/r02_bench_train.log | cut -c1-300
du -sh gpurun_out
