#!/bin/bash
# Second GPU visit of round 2: full test suite, cfg3, A/B of the gather variants, launch list + full ncu capture of the hot kernels.
TAG=r02b
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -15
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --workload cfg3 > gpurun_out/${TAG}_bench_cfg3.json 2> gpurun_out/${TAG}_bench_cfg3.err; tail -c 2500 gpurun_out/${TAG}_bench_cfg3.json; tail -5 gpurun_out/${TAG}_bench_cfg3.err
timeout 600 python profiles/ab_gather2.py > gpurun_out/${TAG}_ab2.txt 2>&1; cat gpurun_out/${TAG}_ab2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --mode static --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/${TAG}_launches.csv 4 > gpurun_out/${TAG}_launches_summary.md 2>&1
head -50 gpurun_out/${TAG}_launches_summary.md
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_fused_sdf_tc|k_sdf_bwd_tc|k_color_' -c 10 -f -o gpurun_out/${TAG}_hot \
  python bench.py --mode static --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_ncu_hot.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_hot.log; ls -la gpurun_out/${TAG}_hot.ncu-rep
