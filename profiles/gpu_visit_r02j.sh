#!/bin/bash
# r02j: recorded one-pass march (small batches by default; forced at frame scale for the A/B) and the chunked hit-list search of
# nsb_assemble_boundary: full GPU suite, smoke, A/B in one process, the driver-style bench line (default and with the one-pass march forced).
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" gpurun_out/${TAG}_pytest.log | tail -8
NSB_MARCH_ONEPASS=1 timeout 600 python -m pytest tests/test_static_gpu.py tests/test_frame_parity_gpu.py tests/test_cfg3_gpu.py -m gpu -q > gpurun_out/${TAG}_pytest_onepass.log 2>&1
echo "pytest(onepass forced) exit $?" >> gpurun_out/${TAG}_pytest_onepass.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest" gpurun_out/${TAG}_pytest_onepass.log | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | cut -c1-300
timeout 300 python profiles/ab_glue.py --steps 20 --rounds 2 > gpurun_out/${TAG}_ab_frame.txt 2>&1; tail -7 gpurun_out/${TAG}_ab_frame.txt | cut -c1-260
timeout 300 python profiles/ab_glue.py --rays 4096 --random-rays --steps 100 --warmup 10 --rounds 2 > gpurun_out/${TAG}_ab_4096.txt 2>&1; tail -7 gpurun_out/${TAG}_ab_4096.txt | cut -c1-260
show() { python - "$1" <<'PY'
import json, sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], round(l["value"],3), "Mrays/s e2e", round(l["e2e"]["value"],3), "ms", round(l["ms_per_step"],4), "median", round(l["median"]["ms_per_step"],4),
          l.get("vs_reference_cuda"), l["step_ms"]["resident_stats"], {k: v["rel_l2"] for k, v in (l.get("parity_vs_reference_kernels") or {}).items() if isinstance(v, dict)})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; show gpurun_out/${TAG}_bench.json
NSB_MARCH_ONEPASS=1 timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_onepass.json 2> gpurun_out/${TAG}_bench_onepass.err; show gpurun_out/${TAG}_bench_onepass.json
timeout 600 python bench.py --rays 4096 --random-rays --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err; show gpurun_out/${TAG}_bench_4096.json
