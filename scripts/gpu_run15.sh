set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "render_multi or two_renders" 2>&1 | tail -4
python bench.py --single-process --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_sp2.json 2> gpurun_out/r2_sp2.err; cut -c1-700 gpurun_out/r2_sp2.json; tail -3 gpurun_out/r2_sp2.err
