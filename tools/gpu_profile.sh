#!/bin/bash
# Runs on the GPU box: ncu launch list of one bench step + full capture of one kernel.
set -u
mkdir -p gpurun_out
KREGEX=${KREGEX:-conv1d_kernel}
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-0} -c ${COUNT:-1400} --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-train > gpurun_out/ncu_bench.log 2>&1
echo "ncu list exit $?"; wc -l gpurun_out/launches.csv
if [ "${FULL:-1}" = "1" ]; then
  echo "== ncu full capture of $KREGEX"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s ${FSKIP:-60} -c ${FCOUNT:-3} \
    -o gpurun_out/prof -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-train > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; ls -la gpurun_out/prof.ncu-rep
fi
