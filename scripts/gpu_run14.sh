set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render_multi or two_renders" 2>&1 | tail -4
python bench.py --single-process --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_sp2.json 2> gpurun_out/r2_sp2.err; cut -c1-400 gpurun_out/r2_sp2.json; tail -3 gpurun_out/r2_sp2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_tr2.json 2> gpurun_out/r2_tr2.err; cut -c1-400 gpurun_out/r2_tr2.json; tail -3 gpurun_out/r2_tr2.err
python bench.py --steps 3 --warmup 3 --no-closest-hit --no-cpu-baseline > gpurun_out/r2_n1.json 2> gpurun_out/r2_n1.err; cut -c1-300 gpurun_out/r2_n1.json
