#!/bin/bash
# Seventh GPU visit: full suite on the final code, default bench line, launch list + ncu of the hot kernels, cfg3 + 4096-ray lines.
TAG=r02g
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -15
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-200; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02g_bench.json").read().strip().splitlines()[-1])
print("bench:", l["value"], l["e2e"]["value"], l["ms_per_step"], l["median"]["ms_per_step"], l["launches_per_step"], l["roofline"]["frac"], l.get("vs_reference_cuda"), l["step_ms"]["resident_stats"], l["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; tail -c 500 gpurun_out/${TAG}_bench_reference.json
timeout 600 python bench.py --rays 4096 --random-rays --no-cpu-baseline > gpurun_out/${TAG}_bench_4096.json 2>/dev/null; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02g_bench_4096.json").read().strip().splitlines()[-1])
print("4096:", l["value"], l["ms_per_step"], l["launches_per_step"], l.get("vs_reference_cuda"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --mode static --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv 4 > gpurun_out/${TAG}_launches_summary.md 2>&1; head -12 gpurun_out/${TAG}_launches_summary.md
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_sdf_tc|k_sdf_bwd_tc|k_color_' -c 12 -f -o gpurun_out/${TAG}_hot \
  python bench.py --mode static --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_hot.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_hot.log
