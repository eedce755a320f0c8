# ncu evidence of round 2 (one GPU).  Every profiled command first runs plain (and must exit 0).
set -x
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-parity"
$B > gpurun_out/ncu_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_cfg2.csv $B > gpurun_out/ncu_bench.log 2>&1
$B > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:heatmap_team -s 10 -c 2 -f -o gpurun_out/r02_prof_cfg2 $B > gpurun_out/ncu_full_cfg2.log 2>&1
H="python profiles/bench_heads.py --only 0"
$H > gpurun_out/ncu_plain_heads.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:simdr_heads_kernel -s 3 -c 2 -f -o gpurun_out/r02_prof_heads $H > gpurun_out/ncu_full_heads.log 2>&1
for only in 9 4 6 2; do
  C="python profiles/bench_configs.py --overlap --only $only"
  $C > gpurun_out/ncu_plain_cfg_$only.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:heatmap_team -s 4 -c 2 -f -o gpurun_out/r02_prof_cfgrow_$only $C > gpurun_out/ncu_full_cfgrow_$only.log 2>&1
done
C="python profiles/bench_configs.py --overlap --only 5"
$C > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:simdr_ring -s 4 -c 2 -f -o gpurun_out/r02_prof_simdr $C > gpurun_out/ncu_full_simdr.log 2>&1
K="python profiles/bench_kernels.py"
$K > gpurun_out/ncu_plain_kernels.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"loss_multi|simdr_sl1|simdr_backward|render_simdr|pck_accumulate" -c 40 --csv --log-file gpurun_out/r02_kernels_ncu.csv $K > gpurun_out/ncu_kernels.log 2>&1
ls -la gpurun_out/*.ncu-rep
