set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/ab_stages.py --scene book2 --spp 144 new fused:RT2025_MEDIA_FIRST=2 cap25:RT2025_PATHS_IN_FLIGHT=33554432 cap26:RT2025_PATHS_IN_FLIGHT=67108864 cap23:RT2025_PATHS_IN_FLIGHT=8388608 2>&1 | tee gpurun_out/r2_ab23.log
python scripts/ab_stages.py --scene cornell --spp 144 new cap25:RT2025_PATHS_IN_FLIGHT=33554432 2>&1 | tee -a gpurun_out/r2_ab23.log
python scripts/ab_stages.py --scene final --spp 16 new cap25:RT2025_PATHS_IN_FLIGHT=33554432 2>&1 | tee -a gpurun_out/r2_ab23.log
