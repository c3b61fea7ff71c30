#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/bench_convnext.py 20
for pass in 1 2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:convnext_fused_kernel -s $((5 + pass)) -c 1 \
    -o gpurun_out/r02_prof_cnx$pass -f python tools/bench_convnext.py 2 > gpurun_out/ncu_cnx$pass.log 2>&1
  echo "pass $pass exit $?"
done
ls -la gpurun_out/r02_prof_cnx*.ncu-rep
