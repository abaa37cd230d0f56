set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_tests49.log
python scripts/ab_stages.py --scene book2 --spp 144 compact part:RT2025_SMEM_NODES_KB=150 2>&1 | tee gpurun_out/r2_ab49.log
python scripts/ab_stages.py --scene book2 --spp 961 compact 2>&1 | tee -a gpurun_out/r2_ab49.log
