mkdir -p gpurun_out
RT2025_ITER_TIMES=1 TAG=walk python scripts/part_stages.py 8 2>&1 | tail -24 > gpurun_out/r2_iter56.log
RT2025_ITER_TIMES=1 TAG=nowalk RT2025_WALK_MIN_DEPTH=0 python scripts/part_stages.py 8 2>&1 | tail -43 >> gpurun_out/r2_iter56.log
cat gpurun_out/r2_iter56.log
