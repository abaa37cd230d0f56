set -x
mkdir -p gpurun_out
python scripts/ab_stages.py --scene book2 --spp 144 default park:RT2025_PARK_LEAVES=1 park_novote:lib=librt2025_novote.so:RT2025_PARK_LEAVES=1 wide:RT2025_WIDE_BVH=1 wide_novote:lib=librt2025_novote.so:RT2025_WIDE_BVH=1 2>&1 | tee gpurun_out/r2_ab43.log
python scripts/ab_stages.py --scene cornell --spp 144 default park:RT2025_PARK_LEAVES=1 park_novote:lib=librt2025_novote.so:RT2025_PARK_LEAVES=1 2>&1 | tee -a gpurun_out/r2_ab43.log
python scripts/ab_stages.py --scene final --spp 16 default novote:lib=librt2025_novote.so 2>&1 | tee -a gpurun_out/r2_ab43.log
