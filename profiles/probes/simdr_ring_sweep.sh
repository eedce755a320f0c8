# per-warp ring of the SimDR decode: warps per CTA x stages per warp (LHN_SIMDR_WARPS / LHN_SIMDR_STAGES)
for w in 8 12 16 24; do for st in 2 3; do echo "warps=$w stages=$st"; LHN_SIMDR_WARPS=$w LHN_SIMDR_STAGES=$st python profiles/bench_configs.py --only 5 | tail -1; done; done
