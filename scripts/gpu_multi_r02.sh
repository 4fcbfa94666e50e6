mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L > gpurun_out/r02_gpus_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.log 2>&1; echo "n$N rc=$?"
tail -n 1 gpurun_out/r02_bench_n$N.log | cut -c1-500
if [ "$N" = "2" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --mode train --gpus $N --steps 3 --warmup 1 > gpurun_out/r02_bench_train_n$N.log 2>&1; echo "train n$N rc=$?"
  tail -n 1 gpurun_out/r02_bench_train_n$N.log | cut -c1-700
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 --no-cpu-extras > gpurun_out/r02_bench_ref_n$N.log 2>&1; echo "ref n$N rc=$?"
  tail -n 1 gpurun_out/r02_bench_ref_n$N.log | cut -c1-200
fi
