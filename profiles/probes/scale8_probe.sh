# N = 8: where does the step time go?  (a) NVML sampling period, (b) graph vs eager
for args in "--clock-interval-ms 0.5" "--clock-interval-ms 5" "--clock-interval-ms 5 --no-graph-all"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29590 bench.py --gpus 8 --steps 50 --warmup 10 --no-e2e $args 2>&1 | tail -1 > gpurun_out/s8.json
  python -c "import json; d=json.load(open('gpurun_out/s8.json')); print('$args', '|', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['clocks'])"
done
