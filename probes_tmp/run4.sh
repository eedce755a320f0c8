python -m pytest tests -m gpu -q -x > gpurun_out/t_full.log 2>&1; tail -3 gpurun_out/t_full.log | cut -c1-220
python profiles/bench_kernels.py --json gpurun_out/r02_kernels_b.json > gpurun_out/r02_kernels_b.txt 2>&1; cut -c1-200 gpurun_out/r02_kernels_b.txt
