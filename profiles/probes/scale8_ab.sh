# N = 8 A/B: the timed steps as one CUDA graph (default at N > 1 so far) vs eager launches
for f in "" "--no-graph-all"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 100 --warmup 10 --no-e2e $f 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=8 %-16s %.2f M samples/s  %.2f us/step   %s' % ('$f' or 'graph', d['value']/1e6, d['ms_per_step']*1e3, d['config']['launch'][:50]))"
done
