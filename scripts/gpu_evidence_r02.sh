# Round-2 evidence: tests, smoke, bench arms, ncu launch list of the bench command, DRAM traffic of the conv launches, --set full captures.
# Outputs under gpurun_out/ (scratch); scripts/make_profiles_r02.py turns them into the committed summaries under profiles/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_gpu.txt 2>&1
PT="python -m pytest -m gpu -q -p no:cacheprovider --timeout=600 --timeout-method=thread"
timeout 1800 $PT tests > gpurun_out/r02_t_all.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r02_t_all.log | cut -c1-200
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.log 2>&1; echo "bench rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_bf16.log 2>&1; echo "bench bf16 rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.log 2>&1; echo "bench ref rc=$?"
timeout 900 python bench.py --mode train --steps 3 --warmup 1 > gpurun_out/r02_bench_train.log 2>&1; echo "bench train rc=$?"
timeout 600 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_prof_convs_b64.log 2>&1
timeout 600 python scripts/bench_conditions.py 64 > gpurun_out/r02_conditions.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_umma_kernel|conv_halo_kernel" -c 400 --csv --log-file gpurun_out/r02_conv_traffic.csv env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_ncu_traffic.log 2>&1; echo "traffic rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:conv_umma_kernel -c 400 -o gpurun_out/r02_conv_umma_full -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"conv_halo_kernel|attn_combine" -c 40 -o gpurun_out/r02_attn_full -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"instnorm_apply|hunfold|hfold|replicate_pad|seg_unfold3" -c 120 -o gpurun_out/r02_ops_full -f env HOIG_PROFILE_SINGLE=1 python scripts/profile_convs.py 64 f16 > gpurun_out/r02_ncu_ops.log 2>&1; echo "ncu ops rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"rasterize_kernel|rast_bin" -c 2 -o gpurun_out/r02_rast_full -f python scripts/bench_rasterizer.py 256 > gpurun_out/r02_ncu_rast.log 2>&1; echo "ncu rast rc=$?"
tail -n 1 gpurun_out/r02_bench.log | cut -c1-600; tail -n 1 gpurun_out/r02_bench_ref.log | cut -c1-300; tail -n 1 gpurun_out/r02_bench_train.log | cut -c1-300
du -sh gpurun_out
