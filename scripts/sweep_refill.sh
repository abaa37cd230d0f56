#!/bin/bash
# closest-hit throughput of the config-5 soups against RT2025_REFILL_MIN (idle lanes a warp waits for before refilling)
for n in "$@"; do
  for r in ${REFILLS:-1 8 16 24 32}; do
    echo "== N=$n refill_min=$r"
    RT2025_REFILL_MIN=$r timeout 300 python bench_closest_hit.py --sizes $n --shapes tri_soup --no-oracle 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('  %-60s %8.1f Mrays/s  frac %.2f' % (d['config']['workload'], d['value'], d['roofline']['frac']))"
  done
done
