#!/bin/bash
# Fourth GPU visit: persistent per-ray kernel as the default no-grad half, robust static step, distant model, compose; bench A/B persistent on/off.
TAG=r02d
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -15
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02d_bench.json").read().strip().splitlines()[-1])
print("persistent on :", l["value"], l["e2e"]["value"], l["ms_per_step"], l["launches_per_step"], l["roofline"]["frac"], l["roofline"]["per_kernel_ms_per_step"], l.get("vs_reference_cuda"), l["step_ms"]["resident_stats"])
PY
NSB_PERSISTENT_UPSAMPLE=0 timeout 900 python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_stages.json 2> gpurun_out/${TAG}_bench_stages.err; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02d_bench_stages.json").read().strip().splitlines()[-1])
print("persistent off:", l["value"], l["e2e"]["value"], l["ms_per_step"], l["launches_per_step"], l["roofline"]["per_kernel_ms_per_step"])
PY
timeout 600 python bench.py --rays 4096 --random-rays --no-cpu-baseline > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02d_bench_4096.json").read().strip().splitlines()[-1])
print("4096 rays     :", l["value"], l["ms_per_step"], l["launches_per_step"], l.get("vs_reference_cuda"), l["step_ms"]["resident_stats"])
PY
NSB_PERSISTENT_UPSAMPLE=0 timeout 600 python bench.py --rays 4096 --random-rays --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_4096_stages.json 2>/dev/null; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02d_bench_4096_stages.json").read().strip().splitlines()[-1])
print("4096, stages  :", l["value"], l["ms_per_step"], l["launches_per_step"])
PY
timeout 900 python bench.py --workload cfg3 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02d_bench_cfg3.json").read().strip().splitlines()[-1])
print("cfg3          :", l["value"], l["ms_per_step"], l.get("launches_per_step"), l.get("vs_reference_cuda"))
PY
