for sg in 0 300 600 1000; do
  for ov in "--no-overlap" ""; do
    LHN_STAGGER_NS=$sg python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline $ov | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stagger', $sg, '$ov', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step']*1e3,1), 'us', round(d['roofline']['frac'],4))"
  done
done
