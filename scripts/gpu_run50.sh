set -x
mkdir -p gpurun_out
python scripts/ab_stages.py --scene book2 --spp 144 default rs512:lib=librt2025_rs512.so rs640:lib=librt2025_rs640.so rs768:lib=librt2025_rs768.so b640:lib=librt2025_b640.so 2>&1 | tee gpurun_out/r2_ab50.log
python scripts/ab_stages.py --scene cornell --spp 144 default rs640:lib=librt2025_rs640.so rs768:lib=librt2025_rs768.so 2>&1 | tee -a gpurun_out/r2_ab50.log
