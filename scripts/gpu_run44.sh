set -x
mkdir -p gpurun_out
python scripts/ab_stages.py --scene book2 --spp 144 default wb3:RT2025_WALK_BLOCKS=3 wb2:RT2025_WALK_BLOCKS=2 wb1:RT2025_WALK_BLOCKS=1 wb2_s3:RT2025_WALK_BLOCKS=2:RT2025_SHADE_STREAMS=3 wb3_s3:RT2025_WALK_BLOCKS=3:RT2025_SHADE_STREAMS=3 2>&1 | tee gpurun_out/r2_ab44.log
