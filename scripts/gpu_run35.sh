set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/ab_stages.py --scene book2 --spp 144 new 2>&1 | tee gpurun_out/r2_ab35.log
python bench.py --steps 3 --warmup 3 --no-closest-hit --no-cpu-baseline > gpurun_out/r2_bench35.json 2> gpurun_out/r2_bench35.err; cut -c1-250 gpurun_out/r2_bench35.json
