set -x
mkdir -p gpurun_out
for sc in book2 cornell; do
python scripts/ab_stages.py --scene $sc --spp 144 nopf:lib=librt2025_nopf.so pf m7b:lib=librt2025_m7b.so m7f:lib=librt2025_m7f.so m00:lib=librt2025_m00.so pf_s0:RT2025_SHADE_STREAMS=0 pf_s3:RT2025_SHADE_STREAMS=3 2>&1 | tee -a gpurun_out/r2_ab17.log
done
python scripts/ab_stages.py --scene final --spp 16 nopf:lib=librt2025_nopf.so pf m7b:lib=librt2025_m7b.so m7f:lib=librt2025_m7f.so m00:lib=librt2025_m00.so 2>&1 | tee -a gpurun_out/r2_ab17.log
