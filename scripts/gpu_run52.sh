set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/r2_bench52_8.err | tee gpurun_out/r2_bench52_8.json | python scripts/bench_line.py
python bench.py --single-process --gpus 8 --steps 10 --warmup 3 2>gpurun_out/r2_bench52_sp8.err | tee gpurun_out/r2_bench52_sp8.json | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 2>gpurun_out/r2_bench52_4.err | tee gpurun_out/r2_bench52_4.json | python scripts/bench_line.py
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/r2_bench52_2.err | tee gpurun_out/r2_bench52_2.json | python scripts/bench_line.py
