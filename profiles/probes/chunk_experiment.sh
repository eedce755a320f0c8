# A/B script of the dropped experiment profiles/probes/r01_early_rearm_chunked_loads.patch (apply the patch first:
# LHN_TEAM_CHUNKS only exists with it).  Output of the round-1 run: profiles/r01_early_rearm_chunked_loads.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/ch_tests.log; cat gpurun_out/ch_tests.log
echo "== chunks on (default), overlap" > gpurun_out/ch_scaling.txt
python profiles/probes/batch_scaling.py --overlap >> gpurun_out/ch_scaling.txt 2>&1
echo "== LHN_TEAM_CHUNKS=1 (early re-arm only), overlap" >> gpurun_out/ch_scaling.txt
LHN_TEAM_CHUNKS=1 python profiles/probes/batch_scaling.py --overlap >> gpurun_out/ch_scaling.txt 2>&1
echo "== chunks on (default), no overlap" >> gpurun_out/ch_scaling.txt
python profiles/probes/batch_scaling.py >> gpurun_out/ch_scaling.txt 2>&1
cat gpurun_out/ch_scaling.txt
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/ch_bench.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/ch_bench.json').read().strip().splitlines()[-1]); print('bench chunks on :', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
LHN_TEAM_CHUNKS=1 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/ch_bench1.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/ch_bench1.json').read().strip().splitlines()[-1]); print('bench chunks off:', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
