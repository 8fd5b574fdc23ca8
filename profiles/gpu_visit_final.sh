#!/bin/bash
# Final verification: the full GPU suite, smoke(), and the driver's bench invocation on the committed code.
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" gpurun_out/${TAG}_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-400
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - "$TAG" <<'PY'
import json, sys
l=json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1])
print("bench:", round(l["value"],2), "Mrays/s  e2e", round(l["e2e"]["value"],2), " ms", round(l["ms_per_step"],3), "median", round(l["median"]["ms_per_step"],3), "frac", round(l["roofline"]["frac"],3),
      l.get("vs_reference_cuda"), l["step_ms"]["resident_stats"], "parity", {k: v["rel_l2"] for k, v in l["parity_vs_reference_kernels"].items() if isinstance(v, dict)})
PY
