#!/bin/bash
# usage: sweep_env.sh "VAR=val VAR2=val" "VAR=val2" ...  -> stage times at capacity 2^24
for e in "$@"; do
  echo "== $e"
  env $e timeout 120 python scripts/sweep_capacity.py 16777216 | grep "no_binning=0"
done
