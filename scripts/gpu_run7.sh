set -x
mkdir -p gpurun_out
for sc in book2 book1; do
python scripts/ab_stages.py --scene $sc --spp 144 r1:lib=librt2025_r1.so default direct:RT2025_FIFO_SLOTS=0 fifo64_r16:RT2025_FIFO_SLOTS=64:RT2025_REFILL_MIN=16 2>&1 | tee -a gpurun_out/r2_ab7.log
done
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
