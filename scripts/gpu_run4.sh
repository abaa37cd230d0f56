set -x
mkdir -p gpurun_out
export RT2025_REFILL_MIN=16
python scripts/prof_extend.py book2 144 > gpurun_out/r2_prof4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 3 -c 1 -o gpurun_out/r2_fifo_extend_ss python scripts/prof_extend.py book2 144 > gpurun_out/r2_prof4_ncu.log 2>&1
export RT2025_LIB=$PWD/raytracer-2025_b200/librt2025_r1.so
unset RT2025_REFILL_MIN
python scripts/prof_extend.py book2 144 > gpurun_out/r2_prof4_plain_r1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 3 -c 1 -o gpurun_out/r2_r1_extend_ss python scripts/prof_extend.py book2 144 > gpurun_out/r2_prof4_ncu_r1.log 2>&1
tail -3 gpurun_out/r2_prof4_ncu_r1.log
