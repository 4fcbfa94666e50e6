mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref n2 rc=$?"
timeout 600 python scripts/profile_convs.py 64 bf16 > gpurun_out/prof_convs_b64.log 2>&1; echo "prof rc=$?"
tail -n 3 gpurun_out/bench_n2.log; tail -n 2 gpurun_out/bench_ref_n2.log | cut -c1-300; grep -E "instnorm|hunfold|forward" gpurun_out/prof_convs_b64.log
