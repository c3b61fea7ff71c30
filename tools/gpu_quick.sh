#!/bin/bash
# quick GPU check after a kernel change: conv / e2e parity tests, then the forward bench (no CPU arm, no train)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/quick_pytest.log 2>&1
echo "pytest exit $?"; tail -n 15 gpurun_out/quick_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-train > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
echo "bench exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/quick_bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for k in d.get("top_kernels", [])[:14]:
        print("  %5.2f%% n=%2d %.4f ms %s" % (100 * k["share"], k["launches_per_step"], k["avg_ms"], k["kernel"]))
except Exception as e:
    print("no bench json", e)
PY
tail -n 5 gpurun_out/quick_bench.err
