mkdir -p gpurun_out
(TAG=default python scripts/part_stages.py 8
TAG=nowalk RT2025_WALK_MIN_DEPTH=0 RT2025_GEN_MEDIA=0 python scripts/part_stages.py 8
TAG=cap2^25 RT2025_PATHS_IN_FLIGHT=33554432 python scripts/part_stages.py 8
TAG=cap38.5M RT2025_PATHS_IN_FLIGHT=38500000 python scripts/part_stages.py 8
TAG=cap2^24 RT2025_PATHS_IN_FLIGHT=16777216 python scripts/part_stages.py 8
TAG=notail RT2025_TAIL_PATHS=0 python scripts/part_stages.py 8
TAG=default python scripts/part_stages.py 1) 2>&1 | tee gpurun_out/r2_part53.log
