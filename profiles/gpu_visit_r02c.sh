#!/bin/bash
# Third GPU visit of round 2: full suite (static + graph + cfg3 + compose + adapter + EMA), TMA A/B on the frame bench.
TAG=r02c
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -15
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02c_bench.json").read().strip().splitlines()[-1])
print("TMA on :", l["value"], l["ms_per_step"], l["roofline"]["per_kernel_ms_per_step"], l.get("vs_reference_cuda"))
PY
NSB_COLOR_TMA=0 timeout 900 python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_notma.json 2> gpurun_out/${TAG}_bench_notma.err; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02c_bench_notma.json").read().strip().splitlines()[-1])
print("TMA off:", l["value"], l["ms_per_step"], l["roofline"]["per_kernel_ms_per_step"])
PY
timeout 600 python bench.py --rays 4096 --random-rays --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err; tail -c 600 gpurun_out/${TAG}_bench_4096.json
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_color_rad_bwd|k_color_sdf_bwd|k_sdf_bwd_tc' -c 3 -f -o gpurun_out/${TAG}_bwd \
  python bench.py --mode static --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_bwd.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bwd.log
