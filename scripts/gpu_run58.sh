mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for lib in librt2025.so librt2025_old.so; do
RT2025_LIB=$PWD/raytracer-2025_b200/$lib python bench_closest_hit.py --sizes 1000000 --shapes tri_soup sphere_soup --no-oracle 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$lib', d['config']['workload'], round(d['value'],1), 'Mrays/s')
"
done | tee gpurun_out/r2_ch58.log
python scripts/ab_stages.py --scene cornell --spp 144 new old:lib=librt2025_old.so 2>&1 | tee -a gpurun_out/r2_ch58.log
python scripts/ab_stages.py --scene final --spp 16 new old:lib=librt2025_old.so 2>&1 | tee -a gpurun_out/r2_ch58.log
