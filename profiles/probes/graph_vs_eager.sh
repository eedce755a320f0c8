# N = 1 A/B: eager launches vs one CUDA graph of all K launches, with and without the launch-overlap flag
for f in "" "--graph-all" "--no-overlap" "--graph-all --no-overlap"; do
  python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline $f 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-26s %.1f M samples/s  %.2f us/step   %s' % ('$f' or 'eager', d['value']/1e6, d['ms_per_step']*1e3, d['config']['launch'][:60]))"
done
