set -x
mkdir -p gpurun_out
for sc in book2 cornell book1; do
python scripts/ab_stages.py --scene $sc --spp 144 r1:lib=librt2025_r1.so fifo_drain:RT2025_REFILL_MIN=32 fifo_r24:RT2025_REFILL_MIN=24 fifo_r16:RT2025_REFILL_MIN=16 fifo_r8:RT2025_REFILL_MIN=8 fifo_r1:RT2025_REFILL_MIN=1 fifo32_r16:RT2025_REFILL_MIN=16:RT2025_FIFO_SLOTS=32 2>&1 | tee -a gpurun_out/r2_ab6.log
done
python scripts/ab_stages.py --scene final --spp 16 fifo_drain:RT2025_REFILL_MIN=32 fifo_r16:RT2025_REFILL_MIN=16 fifo_r1:RT2025_REFILL_MIN=1 2>&1 | tee -a gpurun_out/r2_ab6.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench_closest_hit.py --sizes 1000000 --no-oracle 2>&1 | tail -8 | cut -c1-200 | tee gpurun_out/r2_ch6_new.log
