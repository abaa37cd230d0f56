set -x
mkdir -p gpurun_out
( time python bench.py ) 2>gpurun_out/r2_bench51.err | tee gpurun_out/r2_bench51.json | python scripts/bench_line.py
tail -5 gpurun_out/r2_bench51.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) 2>gpurun_out/r2_bench51_ref.err | tee gpurun_out/r2_bench51_ref.json | cut -c1-400
tail -4 gpurun_out/r2_bench51_ref.err
