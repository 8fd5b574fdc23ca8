#!/bin/bash
TAG=r02i
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_color_gpu.py tests/test_render_gpu.py tests/test_static_gpu.py tests/test_training_gpu.py tests/test_cfg3_gpu.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | head
for T in 2 1; do
NSB_COLOR_TMA=$T timeout 300 python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_tma$T.json 2>/dev/null; python - $T <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/r02i_bench_tma{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("color_tma", sys.argv[1], ":", round(l["value"],2), round(l["ms_per_step"],3), round(l["median"]["ms_per_step"],3), l["roofline"]["per_kernel_ms_per_step"])
PY
done
