#!/bin/bash
# One GPU-box visit: parity tests, bench line, torch-profiler kernel table, ncu launch list (shares), one ncu --set full
# capture of the hash-gather kernel.  Usage:  gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh TAG [tests|notests] [ab]'   (ab: A/B the gather variants instead of the ncu --set full capture)
TAG=${1:-rXX}
MODE=${2:-tests}
mkdir -p gpurun_out
if [ "$MODE" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -15 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
timeout 300 python profiles/torch_profile_step.py > gpurun_out/${TAG}_torch_prof.txt 2>&1
head -60 gpurun_out/${TAG}_torch_prof.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches_summary.md 2>&1
head -40 gpurun_out/${TAG}_launches_summary.md
if [ "${3:-}" = "ab" ]; then
  timeout 600 python profiles/ab_gather.py > gpurun_out/${TAG}_ab.txt 2>&1
  cat gpurun_out/${TAG}_ab.txt
  exit 0
fi
if [ "${3:-}" = "abs" ]; then
  timeout 600 python profiles/ab_scatter.py > gpurun_out/${TAG}_ab_scatter.txt 2>&1
  cat gpurun_out/${TAG}_ab_scatter.txt
fi
# one --set full capture of every hot kernel of ONE step (the 8 big launches of the first step)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_sdf_tc|k_sdf_bwd_tc|k_color_' -c 8 -f -o gpurun_out/${TAG}_hot \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_hot.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_hot.log
