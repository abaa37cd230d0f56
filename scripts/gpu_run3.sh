set -x
mkdir -p gpurun_out
export RT2025_REFILL_MIN=16
python scripts/prof_extend.py book2 64 > gpurun_out/r2_prof3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 12 -c 2 -o gpurun_out/r2_fifo_extend python scripts/prof_extend.py book2 64 > gpurun_out/r2_prof3_ncu.log 2>&1
tail -3 gpurun_out/r2_prof3_ncu.log
