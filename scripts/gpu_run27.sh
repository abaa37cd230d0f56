set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
RT2025_TAIL_PATHS=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python scripts/ab_stages.py --scene book2 --spp 144 new 2>&1 | tee gpurun_out/r2_ab27.log
python scripts/ab_stages.py --scene book2 --spp 16 new 2>&1 | tee -a gpurun_out/r2_ab27.log
python bench.py --steps 3 --warmup 3 --no-closest-hit --no-cpu-baseline > gpurun_out/r2_bench27.json 2> gpurun_out/r2_bench27.err; cut -c1-250 gpurun_out/r2_bench27.json
