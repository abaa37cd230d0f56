set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 3 --warmup 3 --no-closest-hit --no-cpu-baseline > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err; tail -2 gpurun_out/r2_bench24.err
RT2025_PATHS_IN_FLIGHT=134217728 python bench.py --steps 3 --warmup 3 --no-closest-hit --no-cpu-baseline > gpurun_out/r2_bench24_27.json 2>> gpurun_out/r2_bench24.err
python scripts/bench_configs.py 2>&1 | tee gpurun_out/r2_configs24.log
