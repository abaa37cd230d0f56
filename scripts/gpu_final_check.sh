set -x
mkdir -p gpurun_out
SECONDS=0
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench default rc=$? wall=${SECONDS}s"
SECONDS=0
python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "bench reference rc=$? wall=${SECONDS}s"; cut -c1-400 gpurun_out/r2_bench_reference.json
