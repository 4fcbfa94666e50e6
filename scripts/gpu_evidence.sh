# Round evidence: tests, bench arms, launch list, ncu captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=420 --timeout-method=thread"
timeout 1500 $PT tests > gpurun_out/t_all.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/t_all.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --dtype f16 --no-cpu-baseline > gpurun_out/bench_f16.log 2>&1; echo "bench f16 rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1
timeout 600 python scripts/bench_rasterizer.py 8192 > gpurun_out/bench_rast.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_umma_kernel|conv_halo_kernel" -s 422 -c 211 --csv --log-file gpurun_out/conv_traffic.csv python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_traffic.log 2>&1; echo "traffic rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_umma_kernel -s 196 -c 8 -o gpurun_out/conv_umma_full -f python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:conv_halo_kernel -s 9 -c 3 -o gpurun_out/conv_halo_full -f python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_halo.log 2>&1; echo "ncu halo rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attn_combine|instnorm_apply" -s 60 -c 4 -o gpurun_out/ops_full -f python scripts/profile_convs.py 64 bf16 > gpurun_out/ncu_ops.log 2>&1; echo "ncu ops rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:rasterize_kernel -c 1 -o gpurun_out/rast_full -f python scripts/bench_rasterizer.py 256 > gpurun_out/ncu_rast.log 2>&1; echo "ncu rast rc=$?"
tail -n 1 gpurun_out/bench.log | cut -c1-1200; tail -n 1 gpurun_out/bench_f16.log | cut -c1-300; tail -n 1 gpurun_out/bench_ref.log | cut -c1-400; tail -n 2 gpurun_out/bench_rast.log | cut -c1-300
du -sh gpurun_out
