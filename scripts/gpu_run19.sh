set -x
mkdir -p gpurun_out
export RT2025_TAIL_PATHS=0
python scripts/prof_extend.py final 16 > gpurun_out/r2_prof19_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_shade -s 15 -c 8 -o gpurun_out/r2_final_shade python scripts/prof_extend.py final 16 > gpurun_out/r2_prof19_ncu.log 2>&1
tail -2 gpurun_out/r2_prof19_ncu.log
