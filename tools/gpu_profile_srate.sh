#!/bin/bash
# ncu --set full captures (with SASS source) of the S-rate conv kernels through tools/microbench.py
set -u
mkdir -p gpurun_out
python tools/microbench.py pw2 k11pro k21 k21_96 pw1dw pw1 --iters 20 > gpurun_out/micro.txt 2>&1
cat gpurun_out/micro.txt
for c in ${CASES:-k11pro k21 pw2}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv1d_umma_kernel -s 3 -c 1 \
    -o gpurun_out/r02_prof_$c -f python tools/microbench.py $c --iters 2 > gpurun_out/ncu_$c.log 2>&1
  echo "$c exit $?"
done
ls -la gpurun_out/*.ncu-rep
