set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
python scripts/ab_stages.py --scene book2 --spp 144 r1:lib=librt2025_r1.so new 2>&1 | tee gpurun_out/r2_ab13.log
python scripts/ab_stages.py --scene book2 --spp 16 r1:lib=librt2025_r1.so new 2>&1 | tee -a gpurun_out/r2_ab13.log
python scripts/ab_stages.py --scene cornell --spp 16 r1:lib=librt2025_r1.so new 2>&1 | tee -a gpurun_out/r2_ab13.log
python scripts/ab_stages.py --scene book1 --spp 10 r1:lib=librt2025_r1.so new 2>&1 | tee -a gpurun_out/r2_ab13.log
