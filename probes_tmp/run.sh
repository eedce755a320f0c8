python -m pytest tests -m gpu -q -x -k "exchange or bench_path or decode" > gpurun_out/t.log 2>&1; tail -4 gpurun_out/t.log | cut -c1-220
for c in 1 0; do
echo "--- world=1 probe, LHN_XCH_COURIER=$c"
LHN_XCH_COURIER=$c timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29511 profiles/probes/xch_timing.py 2>&1 | grep -E "us per step|seq" | head -5
done
echo "--- cfg4 N=1 (no exchange)"
python bench.py --config 4 --steps 60 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
