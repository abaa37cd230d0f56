set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
/usr/bin/time -v python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; grep -E "Elapsed|Maximum resident" gpurun_out/r2_bench_default.err
/usr/bin/time -v python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; grep -E "Elapsed" gpurun_out/r2_bench_reference.err; cut -c1-400 gpurun_out/r2_bench_reference.json
