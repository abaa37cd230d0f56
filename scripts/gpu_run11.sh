set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; tail -3 gpurun_out/r2_bench11.err
python bench.py --single-process --gpus 1 --steps 2 --warmup 3 > gpurun_out/r2_bench11_sp.json 2>> gpurun_out/r2_bench11.err
cat gpurun_out/r2_bench11_sp.json | cut -c1-600
