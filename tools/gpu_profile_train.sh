#!/bin/bash
# Runs on the GPU box: ncu launch list of one training step + full captures of the tcgen05 wgrad and forward kernels.
set -u
mkdir -p gpurun_out
echo "== ncu launch list (one train step, B=32)"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv \
  --log-file gpurun_out/train_launches.csv python tools/bench_train.py --steps 1 --warm 1 > gpurun_out/ncu_train.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/train_launches.csv
echo "== ncu full: wgrad tcgen05 (k11 32->32, S-rate)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_wgrad_umma -s 2 -c 1 \
  -o gpurun_out/prof_wgrad -f python tools/bench_wgrad.py > gpurun_out/ncu_wgrad.log 2>&1
echo "exit $?"
echo "== ncu full: fused pw1 forward (conv1d_umma <4,1>)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1d_umma_kernel -s 3 -c 1 \
  -o gpurun_out/prof_pw1 -f python tools/microbench.py pw1dw --iters 2 > gpurun_out/ncu_pw1.log 2>&1
echo "exit $?"
ls -la gpurun_out/*.ncu-rep
