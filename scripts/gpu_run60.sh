mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/ab_stages.py --scene book2 --spp 144 new prev:lib=librt2025_prev.so new2 prev2:lib=librt2025_prev.so 2>&1 | tee gpurun_out/r2_ab60.log
python scripts/ab_stages.py --scene cornell --spp 144 new prev:lib=librt2025_prev.so 2>&1 | tee -a gpurun_out/r2_ab60.log
