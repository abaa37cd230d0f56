set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/ab_stages.py --scene book2 --spp 144 new 2>&1 | tee gpurun_out/r2_ab16.log
export RT2025_TAIL_PATHS=0
python scripts/prof_extend.py book2 144 > gpurun_out/r2_prof16_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 43 -c 14 -o gpurun_out/r2_v2_iter3 python scripts/prof_extend.py book2 144 > gpurun_out/r2_prof16_ncu.log 2>&1
tail -2 gpurun_out/r2_prof16_ncu.log
