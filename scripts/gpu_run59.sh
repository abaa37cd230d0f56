mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/ab_stages.py --scene book2 --spp 144 new old:lib=librt2025_old.so new2 old2:lib=librt2025_old.so 2>&1 | tee gpurun_out/r2_ab59.log
python scripts/ab_stages.py --scene book2 --spp 961 new old:lib=librt2025_old.so 2>&1 | tee -a gpurun_out/r2_ab59.log
