#!/bin/bash
TAG=r02h
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_static_gpu.py tests/test_cfg3_gpu.py tests/test_ray_upsample_gpu.py tests/test_compose.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | head
timeout 300 python bench.py --rays 4096 --random-rays --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_4096.json 2>/dev/null; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02h_bench_4096.json").read().strip().splitlines()[-1])
print("4096:", l["value"], l["ms_per_step"], l["median"]["ms_per_step"], l["launches_per_step"])
PY
timeout 300 python bench.py --no-cpu-baseline --no-ref-cuda --steps 10 --warmup 5 > gpurun_out/${TAG}_bench.json 2>/dev/null; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02h_bench.json").read().strip().splitlines()[-1])
print("frame:", l["value"], l["ms_per_step"], l["median"]["ms_per_step"], l["e2e"]["value"])
PY
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err; tail -3 gpurun_out/${TAG}_bench_cfg4.err | cut -c1-300; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02h_bench_cfg4.json").read().strip().splitlines()[-1])
print("cfg4:", l["value"], l["ms_per_step"], l["launches_per_step"], l.get("reference_cuda"), l.get("vs_reference_cuda"), l["step_ms"]["resident"])
PY
