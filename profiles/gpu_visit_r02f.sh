#!/bin/bash
# Sixth GPU visit (8 GPUs): colour-kernel gather unroll (kernel timers), frame-parity tests, strong scaling with round-robin rows.
TAG=r02f
mkdir -p gpurun_out; rm -f gpurun_out/frame_parity.jsonl
timeout 900 python -m pytest tests/test_frame_parity_gpu.py tests/test_color_gpu.py tests/test_render_gpu.py tests/test_occ_ema.py -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|^frame|^4096|gradient rel" | cut -c1-700
timeout 600 python bench.py --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print("unroll 4 in the colour kernels:", l["value"], l["ms_per_step"], l["roofline"]["per_kernel_ms_per_step"])
PY
NG=$(nvidia-smi -L | wc -l)
for N in 2 4 8; do
  if [ $N -gt $NG ]; then break; fi
  OUT=gpurun_out/${TAG}_n${N}_strong.json
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 10 \
      --scaling strong --no-ref-cuda --no-cpu-baseline > $OUT 2> gpurun_out/${TAG}_n${N}_strong.err
  python - "$OUT" "$N" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"N={sys.argv[2]} strong (round-robin rows) value {l['value']:8.2f} Mrays/s  e2e {l['e2e']['value']:8.2f}  ms {l['ms_per_step']:.3f}  median ms {l['median']['ms_per_step']:.3f}  rays/gpu {l['config']['rays_per_step_per_gpu']}")
except Exception as e:
    print("N=", sys.argv[2], "FAILED", e); print(open(sys.argv[1].replace('.json','.err')).read()[-600:])
PY
done
