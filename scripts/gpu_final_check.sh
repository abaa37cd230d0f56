# the round-end sequence the driver runs, in one gpurun call: GPU tests, smoke(), the default bench line
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/final_smoke.log
python bench.py 2>gpurun_out/final_bench.err | tee gpurun_out/final_bench.json | python scripts/bench_line.py
