set -x
mkdir -p gpurun_out
V=""
for lm in 1 8 12 16 20 24 32; do for rm in 1 4 8 16; do V="$V L${lm}_R${rm}:RT2025_LEAF_MIN=$lm:RT2025_REFILL_MIN=$rm"; done; done
python scripts/ab_stages.py --scene book2 --spp 144 r1:lib=librt2025_r1.so $V 2>&1 | tee gpurun_out/r2_ab5.log
V=""
for lm in 8 16 24; do for rm in 4 16; do V="$V P_L${lm}_R${rm}:RT2025_LEAF_MIN=$lm:RT2025_REFILL_MIN=$rm:RT2025_PARK_LEAVES=1"; done; done
python scripts/ab_stages.py --scene book2 --spp 144 $V 2>&1 | tee -a gpurun_out/r2_ab5.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
