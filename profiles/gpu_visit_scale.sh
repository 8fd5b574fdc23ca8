#!/bin/bash
# Multi-GPU visit: weak and strong ray-shard scaling of the graph step, N = 1, 2, 4, 8 (as many as the box has).
TAG=${1:-r02s}
NG=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ $N -gt $NG ]; then break; fi
  for SC in weak strong; do
    if [ $N -eq 1 ] && [ $SC = strong ]; then continue; fi
    OUT=gpurun_out/${TAG}_n${N}_${SC}.json
    if [ $N -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 10 --no-ref-cuda --no-cpu-baseline > $OUT 2> gpurun_out/${TAG}_n${N}_${SC}.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 20 --warmup 10 \
        --scaling $SC --no-ref-cuda --no-cpu-baseline > $OUT 2> gpurun_out/${TAG}_n${N}_${SC}.err
    fi
    python - "$OUT" "$N" "$SC" <<'PY'
import json, sys
try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"N={sys.argv[2]} {sys.argv[3]:6s} value {l['value']:8.2f} Mrays/s  e2e {l['e2e']['value']:8.2f}  ms {l['ms_per_step']:.3f}  median ms {l['median']['ms_per_step']:.3f}  rays/gpu {l['config']['rays_per_step_per_gpu']}")
except Exception as e:
    print("N=", sys.argv[2], sys.argv[3], "FAILED", e)
PY
  done
done
tail -5 gpurun_out/${TAG}_n2_weak.err 2>/dev/null | cut -c1-300
