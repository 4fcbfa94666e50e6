# Quick GPU check: every -m gpu test, one profile pass, one bench line.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_quick.sh'
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 1500 $PT tests > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/t_all.log | cut -c1-300
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; head -n 24 gpurun_out/prof_convs_b64.log | cut -c1-130
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-1500
