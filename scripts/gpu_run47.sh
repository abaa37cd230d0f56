set -x
mkdir -p gpurun_out
RT2025_WIDE_BVH=1 RT2025_FIFO_SLOTS=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or same_seed or hits_are" 2>&1 | tail -3
python scripts/ab_stages.py --scene book2 --spp 144 default wide_smem_direct:RT2025_WIDE_BVH=1:RT2025_FIFO_SLOTS=0 wide_smem_fifo32:RT2025_WIDE_BVH=1:RT2025_FIFO_SLOTS=32 wide_l1_direct:RT2025_WIDE_BVH=1:RT2025_FIFO_SLOTS=0:RT2025_SMEM_NODES_KB=0 2>&1 | tee gpurun_out/r2_ab47.log
python scripts/ab_stages.py --scene cornell --spp 144 default wide_smem_direct:RT2025_WIDE_BVH=1:RT2025_FIFO_SLOTS=0 2>&1 | tee -a gpurun_out/r2_ab47.log
python scripts/ab_stages.py --scene book1 --spp 9 default wide_smem_direct:RT2025_WIDE_BVH=1:RT2025_FIFO_SLOTS=0 2>&1 | tee -a gpurun_out/r2_ab47.log
