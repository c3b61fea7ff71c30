#!/bin/bash
# Runs on the GPU box (via gpurun): tests, smoke, bench; logs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu" 
timeout 1500 python -m pytest tests -m gpu -q --tb=short -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"
tail -n 40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
if [ "${SANITIZE:-0}" = "1" ]; then
  echo "== sanitizer (smoke)"
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer.log 2>&1
  echo "sanitizer exit $?"; tail -n 15 gpurun_out/sanitizer.log
fi
echo "== bench"
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 20 gpurun_out/bench.err
