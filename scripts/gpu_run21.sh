set -x
mkdir -p gpurun_out
export RT2025_TAIL_PATHS=0
python scripts/prof_extend.py final 16 > gpurun_out/r2_prof21_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__icc_request_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -s 31 -c 16 --csv --log-file gpurun_out/r2_final_iter2b.csv python scripts/prof_extend.py final 16 > gpurun_out/r2_prof21_ncu.log 2>&1
tail -2 gpurun_out/r2_prof21_ncu.log
