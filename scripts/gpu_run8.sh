set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
for sc in book2 cornell book1; do
python scripts/ab_stages.py --scene $sc --spp 144 r1:lib=librt2025_r1.so default classic:RT2025_MEDIA_FIRST=0 2>&1 | tee -a gpurun_out/r2_ab8.log
done
python scripts/ab_stages.py --scene final --spp 16 default 2>&1 | tee -a gpurun_out/r2_ab8.log
