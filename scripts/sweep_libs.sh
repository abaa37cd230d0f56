#!/bin/bash
# compare builds of the library: RT2025_LIB=<path> for each argument
for lib in "$@"; do
  echo "== $lib"
  RT2025_LIB=$PWD/raytracer-2025_b200/$lib timeout 120 python scripts/sweep.py RT2025_SMEM_NODES_KB 32 200
done
