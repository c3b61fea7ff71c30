#!/bin/bash
# Runs on the GPU box: ncu launch list of one forward step + full captures of the top forward kernels.
set -u
mkdir -p gpurun_out
echo "== ncu launch list (one forward step, B=16)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
  --log-file gpurun_out/fwd_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-train > gpurun_out/ncu_fwd.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/fwd_launches.csv
for c in pw1dw pw2 k11pro; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_umma_kernel -s 3 -c 1 \
    -o gpurun_out/prof_$c -f python tools/microbench.py $c --iters 2 > gpurun_out/ncu_$c.log 2>&1
  echo "$c exit $?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_umma -s 1 -c 1 \
  -o gpurun_out/prof_attn -f python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention_conformer and 803" -p no:cacheprovider > gpurun_out/ncu_attn.log 2>&1
echo "attn exit $?"
ls -la gpurun_out/*.ncu-rep
