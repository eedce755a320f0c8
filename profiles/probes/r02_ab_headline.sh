# A/B of the headline step on ONE box: this tree against the tree before the courier CTA (worktree _old at 24ef8c7)
B="bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline --no-parity"
for i in 1 2 3; do
  for d in . _old; do
    (cd $d && python $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$d', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['kernel_ms_isolated'])")
  done
done
