# config 4 at N = 4, three repetitions of 200 steps (the first single run of 60 steps measured 0.0524 ms per step
# against 0.0423 at N = 2 and 0.0431 at N = 8: repeat before believing either)
N=4
for i in 1 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus $N --config 4 --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>gpurun_out/cfg4_n${N}.err | tail -1 > gpurun_out/r02_bench_cfg4_n${N}_courier_rep$i.json
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_cfg4_n${N}_courier_rep$i.json').read()); print(d['value'], d['ms_per_step'], d['parity']['ok'], d['parity']['counters_equal_monolithic'])"
done
