set -x
mkdir -p gpurun_out
python scripts/prof_extend.py book2 144 > gpurun_out/r2_walk_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_walk_launches.csv python scripts/prof_extend.py book2 144 > gpurun_out/r2_walk_ncu_l.log 2>&1
tail -2 gpurun_out/r2_walk_ncu_l.log
