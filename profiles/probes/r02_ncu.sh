# ncu evidence of round 2 (one GPU).  Every profiled command first runs plain (and must exit 0); each .ncu-rep is
# summarised ON THE BOX (profiles/ncu_summary.py: headline counters, stall reasons, source hot spots) and deleted —
# gpurun brings back at most 64 MiB.
set -x
sumrm() { python profiles/ncu_summary.py gpurun_out/$1.ncu-rep 25 > gpurun_out/$1_ncu_summary.txt 2>&1; rm -f gpurun_out/$1.ncu-rep; }
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity"
$B > gpurun_out/ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_cfg2.csv $B > gpurun_out/ncu_bench.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:heatmap_team -s 10 -c 1 -f -o gpurun_out/r02_prof_cfg2 $B > gpurun_out/ncu_full_cfg2.log 2>&1
sumrm r02_prof_cfg2
H="python profiles/bench_heads.py --only 0"
$H > gpurun_out/ncu_plain_heads.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:simdr_heads_kernel -s 3 -c 1 -f -o gpurun_out/r02_prof_heads $H > gpurun_out/ncu_full_heads.log 2>&1
python profiles/ncu_summary.py gpurun_out/r02_prof_heads.ncu-rep 25 > gpurun_out/r02_prof_heads_ncu_summary.txt 2>&1
for only in 9 4 6; do
  C="python profiles/bench_configs.py --overlap --only $only"
  $C > gpurun_out/ncu_plain_cfg_$only.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:heatmap_team -s 4 -c 1 -f -o gpurun_out/r02_prof_cfgrow_$only $C > gpurun_out/ncu_full_cfgrow_$only.log 2>&1
  sumrm r02_prof_cfgrow_$only
done
K="python profiles/bench_kernels.py --json gpurun_out/r02q_kernels.json"
$K > gpurun_out/r02q_kernels.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"loss_multi|simdr_sl1|simdr_backward|render_simdr|pck_accumulate|loss_backward" -c 60 --csv --log-file gpurun_out/r02_kernels_ncu.csv $K > gpurun_out/ncu_kernels.log 2>&1
cat gpurun_out/r02q_kernels.txt
du -sh gpurun_out
