set -x
mkdir -p gpurun_out
export RT2025_LIB=$PWD/raytracer-2025_b200/librt2025_w128x4.so
python scripts/prof_extend.py book2 64 > gpurun_out/r2_walk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 1 -c 1 -o gpurun_out/r2_walk_src python scripts/prof_extend.py book2 64 > gpurun_out/r2_walk_ncu.log 2>&1
tail -2 gpurun_out/r2_walk_ncu.log
