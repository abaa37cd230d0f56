set -x
mkdir -p gpurun_out
for sc in book2 cornell book1; do
python scripts/ab_stages.py --scene $sc --spp 144 r1:lib=librt2025_r1.so fifo_drain:RT2025_REFILL_MIN=32 fifo_r1:RT2025_REFILL_MIN=1 fifo_r8:RT2025_REFILL_MIN=8 fifo_r16:RT2025_REFILL_MIN=16 fifo_r24:RT2025_REFILL_MIN=24 park_r1:RT2025_REFILL_MIN=1:RT2025_PARK_LEAVES=1 park_r16:RT2025_REFILL_MIN=16:RT2025_PARK_LEAVES=1 park_drain:RT2025_REFILL_MIN=32:RT2025_PARK_LEAVES=1 2>&1 | tee -a gpurun_out/r2_ab2.log
done
python scripts/ab_stages.py --scene final --spp 16 fifo_drain:RT2025_REFILL_MIN=32 fifo_r1:RT2025_REFILL_MIN=1 fifo_r16:RT2025_REFILL_MIN=16 park0_r1:RT2025_REFILL_MIN=1:RT2025_PARK_LEAVES=0 park0_drain:RT2025_REFILL_MIN=32:RT2025_PARK_LEAVES=0 2>&1 | tee -a gpurun_out/r2_ab2.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests2.log; tail -5 gpurun_out/r2_tests2.log
