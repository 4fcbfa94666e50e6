mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "n$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref n$N rc=$?"
tail -n 1 gpurun_out/bench_n$N.log | cut -c1-900; tail -n 1 gpurun_out/bench_ref_n$N.log | cut -c1-300
