set -x
run() { n=$1; tag=$2; shift 2; timeout 400 python bench.py "$@" --gpus $n --steps 50 --warmup 10 2> gpurun_out/r02t_n${n}_$tag.err | grep "^{" > gpurun_out/r02t_n${n}_$tag.json; echo "n$n $tag rc=$?"; }
run 2 cfg5 --config 5
run 4 cfg5 --config 5
run 4 cfg4 --config 4
run 2 cfg4 --config 4
run 4 cfg2 --config 2
run 2 cfg2 --config 2
run 4 cfg3 --config 3 --no-e2e
run 4 cfg1 --config 1 --no-e2e
# ncu of the two kernels still far below the roofline (single GPU): plain run first
K="python profiles/bench_kernels.py"
$K > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"simdr_sl1_kernel|pck_accumulate_kernel" -s 2 -c 2 -f -o gpurun_out/r02_prof_weak $K > gpurun_out/ncu_weak.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02_prof_weak.ncu-rep 30 > gpurun_out/r02_prof_weak_ncu_summary.txt 2>&1; rm -f gpurun_out/r02_prof_weak.ncu-rep
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02t_n*_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]; p=d.get("parity",{}); e=d.get("e2e",{})
        print(f.split("/")[-1], "value=%.4g ms/step=%.4f kernel_ms=%.4f parity=%s mono=%s e2e=%.4g" % (d["value"], d["ms_per_step"], r["kernel_ms"], p.get("ok"), p.get("counters_equal_monolithic"), e.get("value",0)))
    except Exception as ex: print(f, "ERR", ex)
PY
