# bf16 + flip headline shape: is the single sweeper warp of a 2-warp team the limit?  Generic (run-time-size) kernel with
# 12 teams of 2 warps against 6 teams of 4 warps (LHN_TEAM_POLICY=0), and the compile-time 2-warp instantiation.
echo "--- fast path (compile-time 64x64, 12 teams x 2 warps)"
python profiles/bench_configs.py --overlap --only 4 2>&1 | grep bf16
echo "--- generic path, 12 teams x 2 warps"
LHN_NO_FAST=1 python profiles/bench_configs.py --overlap --only 4 2>&1 | grep bf16
echo "--- generic path, 6 teams x 4 warps (LHN_TEAM_POLICY=0)"
LHN_NO_FAST=1 LHN_TEAM_POLICY=0 python profiles/bench_configs.py --overlap --only 4 2>&1 | grep bf16
