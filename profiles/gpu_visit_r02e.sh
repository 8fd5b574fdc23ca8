#!/bin/bash
# Fifth GPU visit: full suite after the persistent-kernel auto switch, batched accel, conditioned kwargs, scan tickets; gradient parity errors; bench.
TAG=r02e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -15
cat gpurun_out/grad_parity_errors.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-300; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02e_bench.json").read().strip().splitlines()[-1])
print("bench:", l["value"], l["e2e"]["value"], l["ms_per_step"], l["launches_per_step"], l["roofline"]["frac"], l.get("vs_reference_cuda"), l["step_ms"]["resident_stats"], l.get("cpu_baseline"))
PY
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
