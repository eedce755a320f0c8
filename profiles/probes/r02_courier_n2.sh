python -m pytest tests/test_gpu_exchange.py -m gpu -q -x > gpurun_out/t2.log 2>&1; tail -3 gpurun_out/t2.log | cut -c1-220
N=${N:-2}
for c in 1 0; do
echo "--- world=$N probe, LHN_XCH_COURIER=$c"
LHN_XCH_COURIER=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/probes/xch_timing.py 2>&1 | grep -E "us per step" | head -8
done
echo "--- cfg4 N=$N nvlink"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 4 --steps 60 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench_cfg4_n${N}_courier.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg4_n${N}_courier.json').read()); print(d['value'], d['ms_per_step'], d.get('parity'))"
