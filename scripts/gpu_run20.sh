set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python scripts/ab_stages.py --scene final --spp 16 new 2>&1 | tee gpurun_out/r2_ab20.log
python scripts/ab_stages.py --scene book2 --spp 144 new 2>&1 | tee -a gpurun_out/r2_ab20.log
