for sp in 0 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$sp bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e --spare-sms $sp 2>&1 | tail -1 > gpurun_out/scale8_sp$sp.json
  python -c "import sys,json; d=json.load(open('gpurun_out/scale8_sp$sp.json')); print('spare', $sp, d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
