set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final3_tests.log; cat gpurun_out/final3_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; tail -c 1500 gpurun_out/bench_final3.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final3.json 2>/dev/null; tail -c 600 gpurun_out/bench_ref_final3.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final3.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l_final3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heatmap_team -c 1 -f -o gpurun_out/prof_r01_final3 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_final3.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
