#!/bin/bash
mkdir -p gpurun_out
cd $GRAFT_REPO_ROOT
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 2 -c 1 -o /tmp/mlp128 -f python scripts/one_conv.py mlp128 1 > gpurun_out/ncu_mlp.log 2>&1
ncu -i /tmp/mlp128.ncu-rep --page raw --csv > gpurun_out/mlp128_raw.csv 2>/dev/null
ncu -i /tmp/mlp128.ncu-rep --page source --csv > gpurun_out/mlp128_source.csv 2>/dev/null
ls -la gpurun_out/mlp128_*; tail -3 gpurun_out/ncu_mlp.log
