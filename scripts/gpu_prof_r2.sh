set -x
mkdir -p gpurun_out
python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_final_launches.csv python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_ncu_l.log 2>&1 && \
ncu --set full --clock-control none -s 29 -c 14 -o gpurun_out/r2_final_iter2 python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_ncu_f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 2 -c 1 -o gpurun_out/r2_final_extend_src python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_ncu_e.log 2>&1
tail -1 gpurun_out/r2_final_ncu_e.log
ls -la gpurun_out
