mkdir -p gpurun_out
(TAG=off python scripts/part_stages.py 8
for q in 4096 32768 262144; do for st in 4 8; do TAG=q${q}_s$st RT2025_WALK_DRAIN_QUEUE=$q RT2025_WALK_DRAIN_STEPS=$st python scripts/part_stages.py 8; done; done
TAG=q32768_s6_full RT2025_WALK_DRAIN_QUEUE=32768 python scripts/part_stages.py 1) 2>&1 | tee gpurun_out/r2_part55.log
