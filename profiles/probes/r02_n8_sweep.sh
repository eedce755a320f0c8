set -x
run() { tag=$1; shift; timeout 400 python bench.py "$@" --gpus 8 --steps 50 --warmup 10 2> gpurun_out/r02o_n8_$tag.err | grep "^{" > gpurun_out/r02o_n8_$tag.json; echo "$tag rc=$? $(wc -c < gpurun_out/r02o_n8_$tag.json)"; }
run cfg2 --config 2
LHN_TRIGGER=half run cfg4_half --config 4
LHN_TRIGGER=early run cfg4_early --config 4
run cfg4 --config 4
run cfg4_nccl --config 4 --collective nccl --no-e2e
run cfg5 --config 5
run cfg3 --config 3
run cfg1 --config 1
run cfg2_global --config 2 --global-loss --no-e2e
run cfg2_global_nccl --config 2 --global-loss --collective nccl --no-e2e
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29577 tests/mp_exchange_worker.py 2>&1 | grep -v "^NCCL\|OMP_NUM\|\*\*\*" | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29578 profiles/probes/xch_timing.py 2>/dev/null | grep -v "^NCCL" > gpurun_out/r02o_xch_timing_n8.txt; head -12 gpurun_out/r02o_xch_timing_n8.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02o_n8_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; p=d.get("parity",{}); e=d.get("e2e",{})
        print(f.split("/")[-1], "value=%.4g ms/step=%.4f kernel_ms=%.4f parity=%s mono=%s to=%s e2e=%.4g" % (d["value"], d["ms_per_step"], r["kernel_ms"], p.get("ok"), p.get("counters_equal_monolithic"), p.get("exchange_timeouts"), e.get("value",0)))
    except Exception as ex: print(f, "ERR", ex)
PY
