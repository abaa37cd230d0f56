set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2_walk_tests.log
python scripts/ab_stages.py --scene book2 --spp 144 nowalk:RT2025_WALK_MIN_DEPTH=0 walk256x1 w128x3:lib=librt2025_w128x3.so w128x4:lib=librt2025_w128x4.so 2>&1 | tee gpurun_out/r2_walk_ab.log
python scripts/ab_stages.py --scene cornell --spp 144 nowalk:RT2025_WALK_MIN_DEPTH=0 walk256x1 2>&1 | tee -a gpurun_out/r2_walk_ab.log
