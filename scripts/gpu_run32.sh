set -x
mkdir -p gpurun_out
python scripts/ab_stages.py --scene book2 --spp 144 b512 b640:lib=librt2025_b640.so b384:lib=librt2025_b384.so 2>&1 | tee -a gpurun_out/r2_ab32.log
python scripts/ab_stages.py --scene cornell --spp 144 b512 b640:lib=librt2025_b640.so b384:lib=librt2025_b384.so 2>&1 | tee -a gpurun_out/r2_ab32.log
python scripts/ab_stages.py --scene final --spp 16 b512 b640:lib=librt2025_b640.so b384:lib=librt2025_b384.so 2>&1 | tee -a gpurun_out/r2_ab32.log
