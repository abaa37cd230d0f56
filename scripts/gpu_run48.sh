set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_tests48.log
CFGS=1,2,3,4 NO_ORACLE=1 python scripts/bench_configs.py 2>&1 | tee gpurun_out/r2_configs48.log | tail -12
