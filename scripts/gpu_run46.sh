set -x
mkdir -p gpurun_out
python scripts/ab_stages.py --scene book2 --spp 961 gen3 gen2:lib=librt2025_gen2.so gen4:lib=librt2025_gen4.so nowalk:RT2025_WALK_MIN_DEPTH=0:RT2025_GEN_MEDIA=0 2>&1 | tee gpurun_out/r2_ab46.log
