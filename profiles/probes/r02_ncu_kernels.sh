# ncu evidence for the round-2 rewrites of the off-headline kernels (one GPU): a launch list (durations + DRAM bytes) and
# one --set full capture each, summarised on the box (gpurun brings back at most 64 MiB) and deleted.
export LHN_BENCH_GRAPH=0
K="python profiles/bench_kernels.py"
$K > gpurun_out/ncu_plain_kernels.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"loss_multi|simdr_sl1|simdr_loss_finalize|simdr_backward|render_simdr|pck_accumulate|loss_backward" -c 400 --csv --log-file gpurun_out/r02_kernels_ncu_b.csv $K > gpurun_out/ncu_kernels.log 2>&1
for k in simdr_sl1_joint_kernel render_simdr_kernel pck_accumulate_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -f -o gpurun_out/r02b_$k $K > gpurun_out/ncu_full_$k.log 2>&1
  python profiles/ncu_summary.py gpurun_out/r02b_$k.ncu-rep 20 > gpurun_out/r02b_${k}_ncu_summary.txt 2>&1; rm -f gpurun_out/r02b_$k.ncu-rep
done
du -sh gpurun_out; head -30 gpurun_out/r02b_simdr_sl1_joint_kernel_ncu_summary.txt | cut -c1-160
