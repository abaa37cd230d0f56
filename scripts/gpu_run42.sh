set -x
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_bench42.err | tee gpurun_out/r2_bench42.json | python scripts/bench_line.py
