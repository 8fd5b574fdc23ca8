for f in clocks.sm clocks.max.sm power.draw clocks_event_reasons.active clocks_event_reasons.hw_slowdown clocks_event_reasons.sw_power_cap temperature.gpu; do
  s=$(date +%s%N); nvidia-smi --query-gpu=$f --format=csv,noheader -i 0 > /dev/null; e=$(date +%s%N); echo "$f $(( (e-s)/1000000 )) ms"
done
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda --clock-period-ms $1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['step_ms']['resident']+d['step_ms']['e2e']; print('period', '$1', '$2', round(d['ms_per_step'],2), round(d['e2e']['ms_per_step'],2), 'max step', max(r), 'n>20ms', sum(x>20 for x in r))"; }
run 100; run 100; run 100; run 0; run 0


