#!/bin/bash
# quick GPU check: parity tests + one short bench line (value, ms/step, stage times)
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('Mpaths/s %.1f  ms/step %.1f  stages %s  e2e %.1f  clocks %s' % (d['value']/1e6, d['ms_per_step'], {k:round(v,1) for k,v in d['roofline']['stage_ms_per_step'].items()}, d['e2e']['value']/1e6, d['clocks']))"
