python -m pytest tests -m gpu -q 2>&1 | tail -2 > gpurun_out/final4_tests.log; cat gpurun_out/final4_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_final4.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
for s in 600 0 600 0; do LHN_STAGGER_NS=$s python bench.py --steps 200 --warmup 10 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stagger $s:', d['ms_per_step']*1e3, 'us/step; kernel-only pass', d['roofline']['kernel_ms']*1e3)"; done
