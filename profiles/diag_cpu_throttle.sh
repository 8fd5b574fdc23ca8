# are the random 50-200 ms stalls of the 15 ms step CPU-bandwidth throttling of the container (cgroup cpu.max)?
echo "cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)  nproc: $(nproc)"
stat() { grep -E "nr_periods|nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo; }
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --clock-period-ms $1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['step_ms']['resident']+d['step_ms']['e2e']; print('sampler', '$1', 'OMP', '$OMP_NUM_THREADS', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'max step', max(r), 'n>20ms', sum(x>20 for x in r))"; }
stat; run 0; stat; run 0; stat
export OMP_NUM_THREADS=1 MKL_NUM_THREADS=1
run 0; stat; run 0; stat; run 100; stat; run 100; stat
