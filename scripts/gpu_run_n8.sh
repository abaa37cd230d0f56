set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_tr8.json 2> gpurun_out/r2_tr8.err; cut -c1-300 gpurun_out/r2_tr8.json; tail -2 gpurun_out/r2_tr8.err
python bench.py --single-process --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_sp8.json 2> gpurun_out/r2_sp8.err; cut -c1-300 gpurun_out/r2_sp8.json; tail -2 gpurun_out/r2_sp8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_tr4.json 2> gpurun_out/r2_tr4.err; cut -c1-300 gpurun_out/r2_tr4.json
