# Round-2 final validation: every GPU test, smoke, the default bench line and the bf16 bench line.
mkdir -p gpurun_out
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 900 $PT tests > gpurun_out/r02_t_all.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r02_t_all.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
timeout 200 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_prof_convs_b64.log 2>&1
tail -n 1 gpurun_out/r02_bench.log | cut -c1-400
