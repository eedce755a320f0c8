N=${N:-2}
for t in early half; do
echo "--- world=$N probe courier, LHN_TRIGGER=$t"
LHN_TRIGGER=$t timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/probes/xch_timing.py 2>&1 | grep -E "us per step" | head -8
echo "--- cfg4 N=$N nvlink LHN_TRIGGER=$t"
LHN_TRIGGER=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 4 --steps 60 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench_cfg4_n${N}_courier_$t.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg4_n${N}_courier_$t.json').read()); print(d['value'], d['ms_per_step'], d['parity']['ok'], d['parity']['counters_equal_monolithic'])"
done
