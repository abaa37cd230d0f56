set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
python scripts/ab_stages.py --scene book2 --spp 144 r1:lib=librt2025_r1.so pre:RT2025_MEDIA_FIRST=1 fused:RT2025_MEDIA_FIRST=2 classic:RT2025_MEDIA_FIRST=0 2>&1 | tee -a gpurun_out/r2_ab9.log
python scripts/ab_stages.py --scene final --spp 16 default 2>&1 | tee -a gpurun_out/r2_ab9.log
