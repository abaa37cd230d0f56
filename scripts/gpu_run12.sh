set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
RT2025_TAIL_PATHS=0 timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/ab_stages.py --scene book2 --spp 144 notail:RT2025_TAIL_PATHS=0 tail64k tail256k:RT2025_TAIL_PATHS=262144 tail16k:RT2025_TAIL_PATHS=16384 2>&1 | tee gpurun_out/r2_ab12.log
python scripts/ab_stages.py --scene book2 --spp 16 notail:RT2025_TAIL_PATHS=0 tail64k tail256k:RT2025_TAIL_PATHS=262144 tail1M:RT2025_TAIL_PATHS=1048576 2>&1 | tee -a gpurun_out/r2_ab12.log
python scripts/ab_stages.py --scene cornell --spp 16 notail:RT2025_TAIL_PATHS=0 tail64k tail256k:RT2025_TAIL_PATHS=262144 2>&1 | tee -a gpurun_out/r2_ab12.log
python scripts/ab_stages.py --scene final --spp 4 notail:RT2025_TAIL_PATHS=0 tail64k tail256k:RT2025_TAIL_PATHS=262144 2>&1 | tee -a gpurun_out/r2_ab12.log
