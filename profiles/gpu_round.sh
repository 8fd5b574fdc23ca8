#!/bin/bash
# One GPU-box visit.  Usage:  gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh TAG [tests|notests] [flags]'
#   flags (any of, concatenated):  ab = A/B of the SDF query variants, abs = scatter microbenchmark, prof = torch profile + ncu launch list,
#                                  ncu = one `ncu --set full` capture of every hot kernel of one step
TAG=${1:-rXX}
MODE=${2:-tests}
FLAGS=${3:-prof}
mkdir -p gpurun_out
if [ "$MODE" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -8 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json
if [[ "$FLAGS" == *prof* ]]; then
  timeout 300 python profiles/torch_profile_step.py > gpurun_out/${TAG}_torch_prof.txt 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_bench.log 2>&1
  python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv 4 > gpurun_out/${TAG}_launches_summary.md 2>&1
  head -36 gpurun_out/${TAG}_launches_summary.md
fi
if [[ "$FLAGS" == *abs* ]]; then
  timeout 600 python profiles/ab_scatter.py > gpurun_out/${TAG}_ab_scatter.txt 2>&1
  cat gpurun_out/${TAG}_ab_scatter.txt
elif [[ "$FLAGS" == *ab* ]]; then
  timeout 600 python profiles/ab_gather.py > gpurun_out/${TAG}_ab.txt 2>&1
  cat gpurun_out/${TAG}_ab.txt
fi
if [[ "$FLAGS" == *ncu* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_sdf_tc|k_sdf_bwd_tc|k_color_' -c 8 -f -o gpurun_out/${TAG}_hot \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_hot.log 2>&1
  tail -3 gpurun_out/${TAG}_ncu_hot.log
fi
