N=${N:-8}
echo "--- cfg4 N=$N nvlink (courier CTA, early trigger)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 4 --steps 60 --warmup 5 2>gpurun_out/cfg4_n${N}.err | tail -1 > gpurun_out/r02_bench_cfg4_n${N}_courier.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg4_n${N}_courier.json').read()); print(d['value'], d['ms_per_step'], d['parity']['ok'], d['parity']['counters_equal_monolithic'], d['clocks'])"
echo "--- world=$N probe"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/probes/xch_timing.py > gpurun_out/r02_xch_timing_n${N}_courier.txt 2>&1; grep -E "us per step" gpurun_out/r02_xch_timing_n${N}_courier.txt | head -8
echo "--- cfg4 N=1 same box"
python bench.py --config 4 --steps 60 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r02_bench_cfg4_n1_samebox.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg4_n1_samebox.json').read()); print(d['value'], d['ms_per_step'])"
