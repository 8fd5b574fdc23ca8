#!/bin/bash
# First GPU visit of round 2: the static / graph step against the host-sized path, then the standing evidence.
TAG=r02a
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_static_gpu.py -x -q > gpurun_out/${TAG}_static.log 2>&1; echo "static exit $?" >> gpurun_out/${TAG}_static.log
tail -25 gpurun_out/${TAG}_static.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_static_gpu.py --durations=5 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
NSB_TEST_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_experiments_gpu.py -x -q > gpurun_out/${TAG}_exp.log 2>&1; tail -5 gpurun_out/${TAG}_exp.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_graph.json 2> gpurun_out/${TAG}_bench_graph.err; tail -c 2500 gpurun_out/${TAG}_bench_graph.json; tail -5 gpurun_out/${TAG}_bench_graph.err
timeout 600 python bench.py --mode host --no-ref-cuda --no-cpu-baseline > gpurun_out/${TAG}_bench_host.json 2> gpurun_out/${TAG}_bench_host.err; tail -c 1200 gpurun_out/${TAG}_bench_host.json
timeout 600 python bench.py --rays 4096 --random-rays --no-cpu-baseline > gpurun_out/${TAG}_bench_4096.json 2> gpurun_out/${TAG}_bench_4096.err; tail -c 1500 gpurun_out/${TAG}_bench_4096.json; tail -3 gpurun_out/${TAG}_bench_4096.err
timeout 600 python profiles/ab_gather.py > gpurun_out/${TAG}_ab.txt 2>&1; cat gpurun_out/${TAG}_ab.txt
