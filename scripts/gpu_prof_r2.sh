# ncu artefacts of the round-2 final build (profiles/r02_v4_*).  One iteration = 13 launches (generate, media sampling, extend,
# bin, walk, 7 shade classes, tail); k_pixel_list precedes the first, so iteration 2 starts at launch 27.
#   bash scripts/gpu_prof_r2.sh A   launch list of a whole render + one whole wavefront iteration with --set full
#   bash scripts/gpu_prof_r2.sh B   k_extend and k_walk with sources           (two calls: gpurun brings back at most 64 MiB)
set -x
mkdir -p gpurun_out
python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_plain.log 2>&1 || exit 1
if [ "$1" = "A" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_final_launches.csv python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_ncu_l.log 2>&1 && \
ncu --set full --clock-control none -s 27 -c 13 -o gpurun_out/r2_final_iter2 python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_ncu_f.log 2>&1
else
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 2 -c 1 -o gpurun_out/r2_final_extend_src python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_ncu_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_walk -s 2 -c 1 -o gpurun_out/r2_final_walk_src python scripts/prof_extend.py book2 400 > gpurun_out/r2_final_ncu_w.log 2>&1
fi
ls -la gpurun_out | tail -8
