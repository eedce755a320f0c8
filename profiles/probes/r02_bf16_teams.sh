# bf16 + flip headline shape (16 KB stages): 12 teams of 2 warps with ONE stage each (default) against 6 teams of 4 warps
# with TWO stages each (LHN_TEAM_POLICY=0), both on the compile-time 64x64 instantiations; generic path for reference.
for i in 1 2; do
echo "--- 12 teams x 2 warps, 1 stage"
python profiles/bench_configs.py --overlap --only 4 2>&1 | grep bf16
echo "--- 6 teams x 4 warps, 2 stages (LHN_TEAM_POLICY=0)"
LHN_TEAM_POLICY=0 python profiles/bench_configs.py --overlap --only 4 2>&1 | grep bf16
done
echo "--- without --overlap: 12x2 / 6x4"
python profiles/bench_configs.py --only 4 2>&1 | grep bf16
LHN_TEAM_POLICY=0 python profiles/bench_configs.py --only 4 2>&1 | grep bf16
