#!/bin/bash
# compare builds of the library: RT2025_LIB=<path> for each argument
for lib in "$@"; do
  echo "== $lib"
  RT2025_LIB=$PWD/raytracer-2025_b200/$lib timeout 120 python scripts/sweep_capacity.py 16777216 | grep "no_binning=0"
done
