set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_tests45.log
python scripts/ab_stages.py --scene book2 --spp 144 nogen:RT2025_GEN_MEDIA=0 gen 2>&1 | tee gpurun_out/r2_ab45.log
python scripts/ab_stages.py --scene book2 --spp 961 nogen:RT2025_GEN_MEDIA=0 gen 2>&1 | tee -a gpurun_out/r2_ab45.log
