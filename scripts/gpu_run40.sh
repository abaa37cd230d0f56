set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_walk_tests.log
python scripts/ab_stages.py --scene book2 --spp 144 nowalk:RT2025_WALK_MIN_DEPTH=0 walk128x4 w128x3:lib=librt2025_w128x3.so w128x6:lib=librt2025_w128x6.so 2>&1 | tee gpurun_out/r2_walk_ab3.log
unset RT2025_LIB
python scripts/prof_extend.py book2 64 > gpurun_out/r2_walk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 1 -c 1 -o gpurun_out/r2_walk_src python scripts/prof_extend.py book2 64 > gpurun_out/r2_walk_ncu.log 2>&1
