# final validation of the round (one GPU): GPU tests, smoke, kernel table, default bench + the reference arm
python -m pytest tests -m gpu -q > gpurun_out/t_final.log 2>&1; tail -3 gpurun_out/t_final.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python profiles/bench_kernels.py --json gpurun_out/r02_kernels_final.json > gpurun_out/r02_kernels_final.txt 2>&1; cut -c1-200 gpurun_out/r02_kernels_final.txt
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-600
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -1 gpurun_out/bench_reference.json | cut -c1-400
