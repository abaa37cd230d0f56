set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests1.log; tail -5 gpurun_out/r2_tests1.log
python scripts/ab_stages.py --scene book2 --spp 144 r1:lib=librt2025_r1.so fifo64 fifo32:RT2025_FIFO_SLOTS=32 fifo64_nocache:RT2025_SMEM_NODES_KB=0 2>&1 | tee gpurun_out/r2_ab1.log
python scripts/ab_stages.py --scene cornell --spp 144 r1:lib=librt2025_r1.so fifo64 2>&1 | tee -a gpurun_out/r2_ab1.log
python scripts/ab_stages.py --scene book1 --spp 64 r1:lib=librt2025_r1.so fifo64 2>&1 | tee -a gpurun_out/r2_ab1.log
python scripts/ab_stages.py --scene final --spp 16 r1:lib=librt2025_r1.so fifo64 2>&1 | tee -a gpurun_out/r2_ab1.log
RT2025_LIB=$PWD/raytracer-2025_b200/librt2025_r1.so python bench_closest_hit.py --sizes 1000000 --no-oracle 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/r2_ch1_r1.log
python bench_closest_hit.py --sizes 1000000 --no-oracle 2>&1 | tail -8 | cut -c1-400 | tee gpurun_out/r2_ch1_new.log
