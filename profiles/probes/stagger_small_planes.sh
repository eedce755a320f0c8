for s in 600 300 150 0; do echo "== LHN_STAGGER_NS=$s"; LHN_STAGGER_NS=$s python profiles/bench_configs.py --only 2,3,4,6,9 | tail -5; done
