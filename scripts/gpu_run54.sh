mkdir -p gpurun_out
(for t in 65536 131072 262144 524288 1048576 2097152; do TAG=tail$t RT2025_TAIL_PATHS=$t python scripts/part_stages.py 8; done
for t in 65536 262144 1048576; do TAG=tail$t RT2025_TAIL_PATHS=$t python scripts/part_stages.py 1; done) 2>&1 | tee gpurun_out/r2_part54.log
